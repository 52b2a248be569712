/*
 * LBAudioDetectiveFingerprint.h — the result type and matcher of the fingerprint path.
 *
 * Drop-in for the reference's LBAudioDetective/LBAudioDetectiveFingerprint.h: same names, argument order,
 * ownership and return conventions.  Each entry point cites the reference declaration (FP.h) and definition
 * (FP.m = LBAudioDetectiveFingerprint.m) it replaces.  The two compare functions run on the GPU (CUDA, sm_100a);
 * there is no CPU fallback — if the CUDA path cannot run they print the reason to stderr and return NaN (never a score computed some
 * other way); LBAudioDetectiveFingerprintCompareToFingerprintStatus returns the status instead.
 *
 * Internal representation: besides the reference's one-byte-per-Boolean arrays (kept so that
 * GetSubfingerprintAtIndex is byte-identical), every subfingerprint is also held packed as two bit planes,
 *   P[w] bit b = Boolean[2*(32w+b)]     ("positive" sign bit of rank 32w+b)
 *   M[w] bit b = Boolean[2*(32w+b)+1]   ("negative" sign bit)
 * with W = LBAudioDetectiveFingerprintPackedWordsPerPlane(L) words per plane, stored P[0..W) then M[0..W).
 */
#ifndef LBAUDIODETECTIVE_FINGERPRINT_H
#define LBAUDIODETECTIVE_FINGERPRINT_H
#include "LBAudioDetectiveTypes.h"
LBAD_EXTERN_C_BEGIN

typedef struct LBAudioDetectiveFingerprint *LBAudioDetectiveFingerprintRef;

/* ---- reference surface ------------------------------------------------------------------------------ */

/* FP.h:27, FP.m:18-26.  Caller owns the result; free with ...Dispose. */
LBAD_API LBAudioDetectiveFingerprintRef LBAudioDetectiveFingerprintNew(UInt32 inSubfingerprintLength);
/* FP.h:35, FP.m:28-39.  NULL is a no-op. */
LBAD_API void LBAudioDetectiveFingerprintDispose(LBAudioDetectiveFingerprintRef inFingerprint);
/* FP.h:45, FP.m:41-59. */
LBAD_API LBAudioDetectiveFingerprintRef LBAudioDetectiveFingerprintCopy(LBAudioDetectiveFingerprintRef inFingerprint);
/* FP.h:59, FP.m:64-66. */
LBAD_API UInt32 LBAudioDetectiveFingerprintGetSubfingerprintLength(LBAudioDetectiveFingerprintRef inFingerprint);
/* FP.h:70, FP.m:68-70. */
LBAD_API UInt32 LBAudioDetectiveFingerprintGetNumberOfSubfingerprints(LBAudioDetectiveFingerprintRef inFingerprint);
/* FP.h:83, FP.m:72-76.  outSubfingerprint: caller-allocated Boolean[L]; returns L. */
LBAD_API UInt32 LBAudioDetectiveFingerprintGetSubfingerprintAtIndex(LBAudioDetectiveFingerprintRef inFingerprint, UInt32 inIndex, Boolean* outSubfingerprint);
/* FP.h:98, FP.m:81-89.  Succeeds only while the fingerprint is empty; otherwise writes the current length back and returns FALSE. */
LBAD_API Boolean LBAudioDetectiveFingerprintSetSubfingerprintLength(LBAudioDetectiveFingerprintRef inFingerprint, UInt32* ioSubfingerprintLength);
/* FP.h:108, FP.m:91-100.  Copies L Booleans. */
LBAD_API void LBAudioDetectiveFingerprintAddSubfingerprint(LBAudioDetectiveFingerprintRef inFingerprint, Boolean* inSubfingerprint);
/* FP.h:122, FP.m:105-117. */
LBAD_API Boolean LBAudioDetectiveFingerprintEqualToFingerprint(LBAudioDetectiveFingerprintRef inFingerprint1, LBAudioDetectiveFingerprintRef inFingerprint2);
/* FP.h:134, FP.m:119-149.  Swap so fp1 has more subfingerprints; max over time offsets of the f32 mean of
 * CompareSubfingerprints; bit-exact with the reference (same integer counts, IEEE divide, same summation order). */
LBAD_API Float32 LBAudioDetectiveFingerprintCompareToFingerprint(LBAudioDetectiveFingerprintRef inFingerprint1, LBAudioDetectiveFingerprintRef inFingerprint2, UInt32 inRange);
/* FP.h:147, FP.m:151-176.  hits/possible over pairs i < MIN(inRange, L(inFingerprint)) step 2; mask from inSubfingerprint1.
 * Both arrays must hold at least MIN(inRange, L) rounded up to even Booleans (the reference reads that many). */
LBAD_API Float32 LBAudioDetectiveFingerprintCompareSubfingerprints(LBAudioDetectiveFingerprintRef inFingerprint, Boolean* inSubfingerprint1, Boolean* inSubfingerprint2, UInt32 inRange);

/* ---- additions (not in the reference): packed access and the reference's only wire format ----------- */

/* LBAudioDetectiveFingerprintCompareToFingerprint with its status spelled out: noErr and *outMatch written, or
 * kLBAudioDetectiveArgumentInvalid (NULL arguments, unsupported or mismatched subfingerprint lengths), kLBAudioDetectiveDeviceUnavailable,
 * kLBAudioDetectiveDeviceError (reason in LBAudioDetectiveSupportLastError()) with *outMatch untouched. */
LBAD_API OSStatus LBAudioDetectiveFingerprintCompareToFingerprintStatus(LBAudioDetectiveFingerprintRef inFingerprint1, LBAudioDetectiveFingerprintRef inFingerprint2, UInt32 inRange, Float32* outMatch);

/* Words per bit plane for subfingerprint length L: 2, 4 or 8 (L <= 128, 256, 512); 0 if L is unsupported. */
LBAD_API UInt32 LBAudioDetectiveFingerprintPackedWordsPerPlane(UInt32 inSubfingerprintLength);
/* Copies 2*W words (P plane then M plane) of subfingerprint inIndex; returns 2*W. */
LBAD_API UInt32 LBAudioDetectiveFingerprintGetPackedSubfingerprintAtIndex(LBAudioDetectiveFingerprintRef inFingerprint, UInt32 inIndex, UInt32* outWords);
/* Appends inCount subfingerprints given packed (2*W words each). */
LBAD_API OSStatus LBAudioDetectiveFingerprintAddPackedSubfingerprints(LBAudioDetectiveFingerprintRef inFingerprint, const UInt32* inWords, UInt32 inCount);
/* '0'/'1' per Boolean, subfingerprints joined by '+' — LBAudioDetectiveTests.m:22-37 (+stringFromFingerprint:).
 * Returns the number of bytes needed including the NUL; writes at most inCapacity bytes. */
LBAD_API size_t LBAudioDetectiveFingerprintToString(LBAudioDetectiveFingerprintRef inFingerprint, char* outString, size_t inCapacity);
/* Inverse of ToString; NULL on malformed input (ragged subfingerprints, characters other than 0/1/+). */
LBAD_API LBAudioDetectiveFingerprintRef LBAudioDetectiveFingerprintFromString(const char* inString);

LBAD_EXTERN_C_END
#endif
