/*
 * LBAudioDetectiveFrame.h — the reference's spectral-image container, same names and signatures
 * (/root/reference/LBAudioDetective/LBAudioDetectiveFrame.h:27-162; upstream calls the type internal, Frame.h:14, but its own test
 * suite drives it directly: LBAudioDetectiveTests.m:158-172).
 *
 * The container lives in host memory, as upstream's does (GetRow hands out a Float32*).  Its two computing functions run on the GPU:
 *   LBAudioDetectiveFrameDecompose            Frame.m:113-153   2-D Haar, rows then columns, any row count and row length
 *   LBAudioDetectiveFrameExtractFingerprint   Frame.m:165-191   signs of the inNumberOfWavelets largest |coefficients| in rank order
 * Both are bit-exact with the reference (true divisions by sqrtf(n) and sqrtf(2); ties keep ascending flat-index order, the stable
 * reading of -sortUsingComparator:).  There is no CPU fallback: without a CUDA device the two functions — void upstream,
 * so there is no status to return — print the reason to stderr and leave the frame / the output untouched
 * (LBAudioDetectiveSupportLastError() holds the message).  NaN coefficients are ordered by their bit pattern (above infinity);
 * upstream's comparator is inconsistent for them.
 *
 * The extraction pipeline does not use this type: it keeps the images on the device (lbad_extract.cu).  This is the drop-in for
 * callers of the Frame API itself.
 */
#ifndef LBAUDIODETECTIVE_FRAME_H
#define LBAUDIODETECTIVE_FRAME_H
#include "LBAudioDetectiveTypes.h"
LBAD_EXTERN_C_BEGIN

typedef struct LBAudioDetectiveFrame *LBAudioDetectiveFrameRef;                                       /* Frame.h:15 */

/* Frame.h:27, Frame.m:22-31 */
LBAD_API LBAudioDetectiveFrameRef LBAudioDetectiveFrameNew(UInt32 inMaxRowCount);
/* Frame.h:35, Frame.m:33-44 (NULL is a no-op) */
LBAD_API void LBAudioDetectiveFrameDispose(LBAudioDetectiveFrameRef inFrame);
/* Frame.h:45, Frame.m:46-62 */
LBAD_API LBAudioDetectiveFrameRef LBAudioDetectiveFrameCopy(LBAudioDetectiveFrameRef inFrame);
/* Frame.h:59, Frame.m:67-69 */
LBAD_API UInt32 LBAudioDetectiveFrameGetNumberOfRows(LBAudioDetectiveFrameRef inFrame);
/* Frame.h:71, Frame.m:71-73: the frame's own storage, valid until the frame is disposed */
LBAD_API Float32* LBAudioDetectiveFrameGetRow(LBAudioDetectiveFrameRef inFrame, UInt32 inRowIndex);
/* Frame.h:83, Frame.m:75-77 */
LBAD_API Float32 LBAudioDetectiveFrameGetValue(LBAudioDetectiveFrameRef inFrame, UInt32 inRowIndex, UInt32 inColumnIndex);
/* Frame.h:93, Frame.m:79-81 */
LBAD_API Boolean LBAudioDetectiveFrameFull(LBAudioDetectiveFrameRef inFrame);
/* Frame.h:110, Frame.m:86-105: copies inCount values; FALSE when the frame is full.  The row length of the frame is the smallest
 * inCount seen.  (Upstream indexes rows[inRowIndex] unchecked; here an index at or beyond the maximum row count returns FALSE.) */
LBAD_API Boolean LBAudioDetectiveFrameSetRow(LBAudioDetectiveFrameRef inFrame, Float32* inRow, UInt32 inRowIndex, UInt32 inCount);
/* Frame.h:121, Frame.m:113-153 — on the GPU */
LBAD_API void LBAudioDetectiveFrameDecompose(LBAudioDetectiveFrameRef inFrame);
/* Frame.h:131 / :141, Frame.m:155-161 */
LBAD_API size_t LBAudioDetectiveFrameFingerprintSize(LBAudioDetectiveFrameRef inFrame);
LBAD_API UInt32 LBAudioDetectiveFrameFingerprintLength(LBAudioDetectiveFrameRef inFrame);
/* Frame.h:151, Frame.m:165-191 — on the GPU.  Sets outFingerprint[2i] for a positive and [2i+1] for a negative coefficient of rank i
 * and, like upstream, writes nothing else: the caller provides 2 * inNumberOfWavelets zeroed Booleans.  (Upstream throws for
 * inNumberOfWavelets beyond rows x rowLength; here the excess ranks are left untouched.) */
LBAD_API void LBAudioDetectiveFrameExtractFingerprint(LBAudioDetectiveFrameRef inFrame, UInt32 inNumberOfWavelets, Boolean* outFingerprint);
/* Frame.h:162, Frame.m:193-210 */
LBAD_API Boolean LBAudioDetectiveFrameEqualToFrame(LBAudioDetectiveFrameRef inFrame1, LBAudioDetectiveFrameRef inFrame2);

/* Additions: the same two operations with a status (kLBAudioDetectiveDeviceUnavailable without a CUDA device). */
LBAD_API OSStatus LBAudioDetectiveFrameDecomposeStatus(LBAudioDetectiveFrameRef inFrame);
LBAD_API OSStatus LBAudioDetectiveFrameExtractFingerprintStatus(LBAudioDetectiveFrameRef inFrame, UInt32 inNumberOfWavelets, Boolean* outFingerprint);

LBAD_EXTERN_C_END
#endif
