/*
 * LBAudioDetectiveFingerprint.c — host side of the result type.  Container operations follow the reference
 * (LBAudioDetectiveFingerprint.m, cited per function); the two compare functions run on the GPU through
 * lbad_search.cu (a one-clip database against a one-query batch) — no arithmetic of the match happens on the CPU.
 */
#include "lbad_host.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

UInt32 lbad_words_per_plane(UInt32 length) {
    UInt32 pairs = (length + 1) / 2;
    if (length == 0 || pairs > 256) return 0;
    return pairs <= 64 ? 2 : pairs <= 128 ? 4 : 8;
}

/* number of (P, M) pairs FP.m:155 visits: i < MIN(range, length), step 2 */
UInt32 lbad_pairs_for_range(UInt32 range, UInt32 length) {
    UInt32 lim = range < length ? range : length;
    return (lim + 1) / 2;
}

void lbad_pack_booleans(const Boolean* in, UInt32 length, UInt32 W, UInt32* out) {
    memset(out, 0, 2 * (size_t)W * sizeof(UInt32));
    for (UInt32 i = 0; i < length; i++) {
        if (!in[i]) continue;
        UInt32 pair = i >> 1;
        out[(i & 1 ? W : 0) + (pair >> 5)] |= 1u << (pair & 31);
    }
}

void lbad_unpack_words(const UInt32* in, UInt32 length, UInt32 W, Boolean* out) {
    for (UInt32 i = 0; i < length; i++) {
        UInt32 pair = i >> 1;
        out[i] = (in[(i & 1 ? W : 0) + (pair >> 5)] >> (pair & 31)) & 1u;
    }
}

OSStatus lbad_status(int code) {
    switch (code) {
        case LBAD_OK: return noErr;
        case LBAD_ERR_ARG: return kLBAudioDetectiveArgumentInvalid;
        case LBAD_ERR_NODEVICE: return kLBAudioDetectiveDeviceUnavailable;
        default: return kLBAudioDetectiveDeviceError;
    }
}

static int reserve(LBAudioDetectiveFingerprintRef fp, UInt32 want) {
    if (want <= fp->capacity) return 1;
    UInt32 cap = fp->capacity ? fp->capacity * 2 : 8;
    if (cap < want) cap = want;
    UInt32 W = lbad_words_per_plane(fp->subfingerprintLength);
    Boolean* b = realloc(fp->booleans, (size_t)cap * (fp->subfingerprintLength ? fp->subfingerprintLength : 1));
    if (!b) return 0;
    fp->booleans = b;
    UInt32* w = realloc(fp->words, (size_t)cap * 2 * (W ? W : 1) * sizeof(UInt32));
    if (!w) return 0;
    fp->words = w;
    fp->capacity = cap;
    return 1;
}

/* FP.m:18-26 */
LBAudioDetectiveFingerprintRef LBAudioDetectiveFingerprintNew(UInt32 inSubfingerprintLength) {
    LBAudioDetectiveFingerprintRef fp = calloc(1, sizeof *fp);
    if (fp) fp->subfingerprintLength = inSubfingerprintLength;
    return fp;
}

/* FP.m:28-39 */
void LBAudioDetectiveFingerprintDispose(LBAudioDetectiveFingerprintRef fp) {
    if (fp == NULL) return;
    free(fp->booleans); free(fp->words); free(fp);
}

/* FP.m:41-59 */
LBAudioDetectiveFingerprintRef LBAudioDetectiveFingerprintCopy(LBAudioDetectiveFingerprintRef in) {
    LBAudioDetectiveFingerprintRef fp = LBAudioDetectiveFingerprintNew(in->subfingerprintLength);
    if (!fp) return NULL;
    if (in->subfingerprintCount) {
        if (!reserve(fp, in->subfingerprintCount)) { LBAudioDetectiveFingerprintDispose(fp); return NULL; }
        UInt32 W = lbad_words_per_plane(in->subfingerprintLength);
        memcpy(fp->booleans, in->booleans, (size_t)in->subfingerprintCount * in->subfingerprintLength);
        memcpy(fp->words, in->words, (size_t)in->subfingerprintCount * 2 * W * sizeof(UInt32));
        fp->subfingerprintCount = in->subfingerprintCount;
    }
    return fp;
}

/* FP.m:64-66 */
UInt32 LBAudioDetectiveFingerprintGetSubfingerprintLength(LBAudioDetectiveFingerprintRef fp) { return fp->subfingerprintLength; }
/* FP.m:68-70 */
UInt32 LBAudioDetectiveFingerprintGetNumberOfSubfingerprints(LBAudioDetectiveFingerprintRef fp) { return fp->subfingerprintCount; }

/* FP.m:72-76 */
UInt32 LBAudioDetectiveFingerprintGetSubfingerprintAtIndex(LBAudioDetectiveFingerprintRef fp, UInt32 inIndex, Boolean* out) {
    memcpy(out, fp->booleans + (size_t)inIndex * fp->subfingerprintLength, fp->subfingerprintLength * sizeof(Boolean));
    return fp->subfingerprintLength;
}

/* FP.m:81-89 */
Boolean LBAudioDetectiveFingerprintSetSubfingerprintLength(LBAudioDetectiveFingerprintRef fp, UInt32* ioLength) {
    if (fp->subfingerprintCount > 0) {
        *ioLength = fp->subfingerprintLength;
        return FALSE;
    }
    fp->subfingerprintLength = *ioLength;
    free(fp->booleans); free(fp->words); fp->booleans = NULL; fp->words = NULL; fp->capacity = 0;
    return TRUE;
}

/* FP.m:91-100 */
void LBAudioDetectiveFingerprintAddSubfingerprint(LBAudioDetectiveFingerprintRef fp, Boolean* inSubfingerprint) {
    if (!reserve(fp, fp->subfingerprintCount + 1)) return;
    UInt32 L = fp->subfingerprintLength, W = lbad_words_per_plane(L);
    memcpy(fp->booleans + (size_t)fp->subfingerprintCount * L, inSubfingerprint, L);
    if (W) lbad_pack_booleans(inSubfingerprint, L, W, fp->words + (size_t)fp->subfingerprintCount * 2 * W);
    fp->subfingerprintCount++;
}

OSStatus lbad_fingerprint_append_packed(LBAudioDetectiveFingerprintRef fp, const UInt32* words, UInt32 count) {
    UInt32 L = fp->subfingerprintLength, W = lbad_words_per_plane(L);
    if (!W) return kLBAudioDetectiveArgumentInvalid;
    if (!reserve(fp, fp->subfingerprintCount + count)) return kLBAudioDetectiveArgumentInvalid;
    for (UInt32 i = 0; i < count; i++) {
        UInt32 at = fp->subfingerprintCount + i;
        memcpy(fp->words + (size_t)at * 2 * W, words + (size_t)i * 2 * W, 2 * W * sizeof(UInt32));
        lbad_unpack_words(words + (size_t)i * 2 * W, L, W, fp->booleans + (size_t)at * L);
    }
    fp->subfingerprintCount += count;
    return noErr;
}

/* FP.m:105-117 */
Boolean LBAudioDetectiveFingerprintEqualToFingerprint(LBAudioDetectiveFingerprintRef a, LBAudioDetectiveFingerprintRef b) {
    if (a->subfingerprintCount != b->subfingerprintCount || a->subfingerprintLength != b->subfingerprintLength) return FALSE;
    if (a->subfingerprintCount == 0) return TRUE;
    return memcmp(a->booleans, b->booleans, (size_t)a->subfingerprintCount * a->subfingerprintLength) == 0;
}

/* The two value-returning compare functions have no status to return an error in (FP.h:134, FP.h:147).  When the CUDA path cannot
 * run — no device, or a CUDA call failed — they say so on stderr, leave the reason in LBAudioDetectiveSupportLastError() and return
 * NaN: loud, never a plausible score and never a silently computed one (there is no CPU fallback), and the host process lives on.
 * LBAudioDetectiveFingerprintCompareToFingerprintStatus is the same call with the status spelled out. */
static Float32 compare_failed(const char* what, int code) {
    fprintf(stderr, "%s: CUDA path unavailable (%d: %s). This library has no CPU fallback; returning NaN.\n", what, code, lbadcu_last_error());
    return (Float32)NAN;
}

/* score = CompareToFingerprint(fp1, fp2, range) on the GPU (compare_pair_kernel, cached per-thread context) */
static int gpu_compare(const UInt32* w1, UInt32 c1, const UInt32* w2, UInt32 c2, UInt32 W, UInt32 pairs, Float32* out) {
    return lbadcu_compare_pair(W, pairs, w1, c1, w2, c2, out);
}

OSStatus LBAudioDetectiveFingerprintCompareToFingerprintStatus(LBAudioDetectiveFingerprintRef fp1, LBAudioDetectiveFingerprintRef fp2, UInt32 inRange, Float32* outMatch) {
    if (!fp1 || !fp2 || !outMatch) return kLBAudioDetectiveArgumentInvalid;
    /* after the swap of FP.m:123-131 the length used by FP.m:155 is that of the fingerprint with more subfingerprints */
    LBAudioDetectiveFingerprintRef longer = fp1->subfingerprintCount < fp2->subfingerprintCount ? fp2 : fp1;
    UInt32 L = longer->subfingerprintLength, W = lbad_words_per_plane(L);
    if (!W || lbad_words_per_plane(fp1->subfingerprintLength) != W || lbad_words_per_plane(fp2->subfingerprintLength) != W) return kLBAudioDetectiveArgumentInvalid;
    Float32 score = 0.0f;
    OSStatus e = lbad_status(gpu_compare(fp1->words, fp1->subfingerprintCount, fp2->words, fp2->subfingerprintCount, W, lbad_pairs_for_range(inRange, L), &score));
    if (e == noErr) *outMatch = score;
    return e;
}

/* FP.m:119-149 */
Float32 LBAudioDetectiveFingerprintCompareToFingerprint(LBAudioDetectiveFingerprintRef fp1, LBAudioDetectiveFingerprintRef fp2, UInt32 inRange) {
    Float32 score = 0.0f;
    OSStatus e = LBAudioDetectiveFingerprintCompareToFingerprintStatus(fp1, fp2, inRange, &score);
    if (e == kLBAudioDetectiveArgumentInvalid) {
        fprintf(stderr, "LBAudioDetectiveFingerprintCompareToFingerprint: unsupported or mismatched subfingerprint lengths (%u, %u)\n",
                fp1 ? (unsigned)fp1->subfingerprintLength : 0u, fp2 ? (unsigned)fp2->subfingerprintLength : 0u);
        return 0.0f;
    }
    if (e != noErr) return compare_failed("LBAudioDetectiveFingerprintCompareToFingerprint", (int)e);
    return score;
}

/* FP.m:151-176 */
Float32 LBAudioDetectiveFingerprintCompareSubfingerprints(LBAudioDetectiveFingerprintRef fp, Boolean* s1, Boolean* s2, UInt32 inRange) {
    UInt32 L = fp->subfingerprintLength, W = lbad_words_per_plane(L);
    if (!W) return 0.0f;
    UInt32 lim = inRange < L ? inRange : L;
    UInt32 n = (lim + 1) & ~1u;                       /* the reference reads pairs, i.e. up to lim rounded up to even */
    UInt32 w1[16], w2[16];
    lbad_pack_booleans(s1, n, W, w1);
    lbad_pack_booleans(s2, n, W, w2);
    Float32 score = 0.0f;
    int e = gpu_compare(w1, 1, w2, 1, W, (lim + 1) / 2, &score);
    if (e != LBAD_OK) return compare_failed("LBAudioDetectiveFingerprintCompareSubfingerprints", e);
    return score;
}

/* ---- additions ---- */

UInt32 LBAudioDetectiveFingerprintPackedWordsPerPlane(UInt32 inSubfingerprintLength) { return lbad_words_per_plane(inSubfingerprintLength); }

UInt32 LBAudioDetectiveFingerprintGetPackedSubfingerprintAtIndex(LBAudioDetectiveFingerprintRef fp, UInt32 inIndex, UInt32* outWords) {
    UInt32 W = lbad_words_per_plane(fp->subfingerprintLength);
    if (!W || inIndex >= fp->subfingerprintCount) return 0;
    memcpy(outWords, fp->words + (size_t)inIndex * 2 * W, 2 * W * sizeof(UInt32));
    return 2 * W;
}

OSStatus LBAudioDetectiveFingerprintAddPackedSubfingerprints(LBAudioDetectiveFingerprintRef fp, const UInt32* inWords, UInt32 inCount) {
    if (!fp || (!inWords && inCount)) return kLBAudioDetectiveArgumentInvalid;
    return lbad_fingerprint_append_packed(fp, inWords, inCount);
}

/* LBAudioDetectiveTests.m:22-37 */
size_t LBAudioDetectiveFingerprintToString(LBAudioDetectiveFingerprintRef fp, char* out, size_t cap) {
    size_t need = (size_t)fp->subfingerprintCount * fp->subfingerprintLength + (fp->subfingerprintCount ? fp->subfingerprintCount - 1 : 0) + 1;
    if (!out || cap == 0) return need;
    size_t o = 0;
    for (UInt32 i = 0; i < fp->subfingerprintCount; i++) {
        if (i && o + 1 < cap) out[o++] = '+';
        for (UInt32 j = 0; j < fp->subfingerprintLength && o + 1 < cap; j++) out[o++] = fp->booleans[(size_t)i * fp->subfingerprintLength + j] ? '1' : '0';
    }
    out[o] = '\0';
    return need;
}

LBAudioDetectiveFingerprintRef LBAudioDetectiveFingerprintFromString(const char* s) {
    if (!s) return NULL;
    size_t first = strcspn(s, "+");
    LBAudioDetectiveFingerprintRef fp = LBAudioDetectiveFingerprintNew((UInt32)first);
    if (!fp) return NULL;
    if (*s == '\0') return fp;
    Boolean* tmp = malloc(first ? first : 1);
    const char* p = s;
    for (;;) {
        size_t n = strcspn(p, "+");
        if (n != first) goto bad;
        for (size_t j = 0; j < n; j++) { if (p[j] != '0' && p[j] != '1') goto bad; tmp[j] = p[j] == '1'; }
        LBAudioDetectiveFingerprintAddSubfingerprint(fp, tmp);
        if (p[n] == '\0') break;
        p += n + 1;
    }
    free(tmp);
    return fp;
bad:
    free(tmp); LBAudioDetectiveFingerprintDispose(fp);
    return NULL;
}
