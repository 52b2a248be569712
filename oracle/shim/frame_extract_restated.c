/* TEST INFRASTRUCTURE — not part of the product.
 *
 * C restatement of the ONE Objective-C function on the path, LBAudioDetectiveFrameExtractFingerprint
 * (LBAudioDetectiveFrame.m:165-191), which cannot be compiled here (no ObjC front-end or runtime in the
 * image).  oracle/Makefile compiles LBAudioDetectiveFrame.m with exactly those lines removed (streamed
 * through sed into gcc, nothing copied) and links this file in their place.  Written against the
 * reference's public Frame getters only.
 *
 * Semantics restated:
 *   Frame.m:168-174  gather rows[r][c] into a flat array at index r*rowLength + c
 *   Frame.m:176-178  sort descending by fabs((double)value).  -[NSMutableArray sortUsingComparator:]
 *                    without NSSortStable is not contractually stable; DEFINED here as stable
 *                    (equal magnitudes keep ascending flat-index order) — SURVEY.md Q9.
 *   Frame.m:182-190  for rank i < inNumberOfWavelets: value > 0 -> out[2i] = 1; value < 0 -> out[2i+1] = 1
 */
#include <Foundation/Foundation.h>
#include "LBAudioDetectiveFrame.h"

static void merge_sort_desc_abs(const Float32* v, UInt32* idx, UInt32* tmp, UInt32 n) {
    for (UInt32 w = 1; w < n; w *= 2) {
        for (UInt32 lo = 0; lo < n; lo += 2 * w) {
            UInt32 mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            UInt32 a = lo, b = mid, o = lo;
            while (a < mid && b < hi) {
                /* take from the right run only if strictly larger: stability */
                if (fabs((double)v[idx[b]]) > fabs((double)v[idx[a]])) tmp[o++] = idx[b++]; else tmp[o++] = idx[a++];
            }
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, n * sizeof(UInt32));
    }
}

void LBAudioDetectiveFrameExtractFingerprint(LBAudioDetectiveFrameRef inFrame, UInt32 inNumberOfWavelets, Boolean* outFingerprint) {
    UInt32 rows = LBAudioDetectiveFrameGetNumberOfRows(inFrame);
    UInt32 rowLength = rows ? LBAudioDetectiveFrameFingerprintLength(inFrame) / (2 * rows) : 0;
    UInt32 n = rows * rowLength;
    Float32* flat = malloc((n ? n : 1) * sizeof(Float32));
    UInt32* idx = malloc((n ? n : 1) * sizeof(UInt32));
    UInt32* tmp = malloc((n ? n : 1) * sizeof(UInt32));
    for (UInt32 r = 0; r < rows; r++) {
        const Float32* row = LBAudioDetectiveFrameGetRow(inFrame, r);
        for (UInt32 c = 0; c < rowLength; c++) { flat[r * rowLength + c] = row[c]; idx[r * rowLength + c] = r * rowLength + c; }
    }
    merge_sort_desc_abs(flat, idx, tmp, n);
    for (UInt32 i = 0; i < inNumberOfWavelets; i++) {
        Float64 value = flat[idx[i]];
        if (value > 0.0) outFingerprint[2 * i] = TRUE;
        else if (value < 0.0) outFingerprint[(2 * i) + 1] = TRUE;
    }
    free(flat); free(idx); free(tmp);
}
