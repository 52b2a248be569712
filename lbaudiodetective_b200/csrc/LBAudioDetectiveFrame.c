/*
 * LBAudioDetectiveFrame.c — the reference's Frame container (LBAudioDetectiveFrame.m) with its two computing functions on the GPU.
 * The container is host memory, as upstream's (GetRow returns the frame's own Float32*); Decompose and ExtractFingerprint gather the
 * rows into one [rows][rowLength] block, run the kernels of lbad_frame.cu on it and scatter the result back.  Declared in
 * include/LBAudioDetectiveFrame.h, every function citing the upstream lines it replaces.
 */
#include "lbad_host.h"
#include "../../include/LBAudioDetectiveFrame.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct LBAudioDetectiveFrame {                 /* Frame.m:11-16 */
    Float32** rows;
    UInt32 maxNumberOfRows;
    UInt32 numberOfRows;
    UInt32 rowLength;
};

LBAudioDetectiveFrameRef LBAudioDetectiveFrameNew(UInt32 inMaxRowCount) {              /* Frame.m:22-31 */
    struct LBAudioDetectiveFrame* f = calloc(1, sizeof *f);
    if (!f) return NULL;
    f->rows = calloc(inMaxRowCount ? inMaxRowCount : 1, sizeof(Float32*));
    if (!f->rows) { free(f); return NULL; }
    f->maxNumberOfRows = inMaxRowCount;
    return f;
}

void LBAudioDetectiveFrameDispose(LBAudioDetectiveFrameRef f) {                        /* Frame.m:33-44 */
    if (!f) return;
    for (UInt32 i = 0; i < f->maxNumberOfRows; i++) free(f->rows[i]);                  /* (upstream frees the first numberOfRows slots; rows set out of order would leak) */
    free(f->rows);
    free(f);
}

LBAudioDetectiveFrameRef LBAudioDetectiveFrameCopy(LBAudioDetectiveFrameRef in) {      /* Frame.m:46-62 */
    if (!in) return NULL;
    LBAudioDetectiveFrameRef f = LBAudioDetectiveFrameNew(in->maxNumberOfRows);
    if (!f) return NULL;
    f->numberOfRows = in->numberOfRows;
    f->rowLength = in->rowLength;
    for (UInt32 i = 0; i < in->maxNumberOfRows; i++) {
        if (!in->rows[i]) continue;
        f->rows[i] = malloc((in->rowLength ? in->rowLength : 1) * sizeof(Float32));    /* rowLength values per row, as upstream */
        if (!f->rows[i]) { LBAudioDetectiveFrameDispose(f); return NULL; }
        memcpy(f->rows[i], in->rows[i], in->rowLength * sizeof(Float32));
    }
    return f;
}

UInt32 LBAudioDetectiveFrameGetNumberOfRows(LBAudioDetectiveFrameRef f) { return f->numberOfRows; }                                  /* Frame.m:67-69 */
Float32* LBAudioDetectiveFrameGetRow(LBAudioDetectiveFrameRef f, UInt32 r) { return f->rows[r]; }                                    /* Frame.m:71-73 */
Float32 LBAudioDetectiveFrameGetValue(LBAudioDetectiveFrameRef f, UInt32 r, UInt32 c) { return f->rows[r][c]; }                      /* Frame.m:75-77 */
Boolean LBAudioDetectiveFrameFull(LBAudioDetectiveFrameRef f) { return f->numberOfRows >= f->maxNumberOfRows; }                      /* Frame.m:79-81 */

Boolean LBAudioDetectiveFrameSetRow(LBAudioDetectiveFrameRef f, Float32* inRow, UInt32 inRowIndex, UInt32 inCount) {                 /* Frame.m:86-105 */
    if (LBAudioDetectiveFrameFull(f)) return FALSE;
    if (inRowIndex >= f->maxNumberOfRows) return FALSE;                                /* (upstream writes out of bounds) */
    Float32* row = calloc(inCount ? inCount : 1, sizeof(Float32));
    if (!row) return FALSE;
    if (inCount) memcpy(row, inRow, inCount * sizeof(Float32));
    free(f->rows[inRowIndex]);                                                         /* (upstream leaks a row that is set twice) */
    f->rows[inRowIndex] = row;
    f->rowLength = f->rowLength == 0 ? inCount : (f->rowLength < inCount ? f->rowLength : inCount);      /* Frame.m:96-101 */
    f->numberOfRows++;
    return TRUE;
}

size_t LBAudioDetectiveFrameFingerprintSize(LBAudioDetectiveFrameRef f) { return (size_t)f->numberOfRows * f->rowLength * 2 * sizeof(Boolean); }     /* Frame.m:155-157 */
UInt32 LBAudioDetectiveFrameFingerprintLength(LBAudioDetectiveFrameRef f) { return f->numberOfRows * f->rowLength * 2; }                            /* Frame.m:159-161 */

/* rows 0 .. numberOfRows-1 as one [numberOfRows][rowLength] block (what both upstream loops walk: Frame.m:114, :168); NULL if a row is missing */
static Float32* gather(LBAudioDetectiveFrameRef f) {
    const size_t n = (size_t)f->numberOfRows * f->rowLength;
    Float32* a = malloc((n ? n : 1) * sizeof(Float32));
    if (!a) return NULL;
    for (UInt32 r = 0; r < f->numberOfRows; r++) {
        if (!f->rows[r]) { free(a); return NULL; }
        memcpy(a + (size_t)r * f->rowLength, f->rows[r], f->rowLength * sizeof(Float32));
    }
    return a;
}

OSStatus LBAudioDetectiveFrameDecomposeStatus(LBAudioDetectiveFrameRef f) {            /* Frame.m:113-132 */
    if (!f) return kLBAudioDetectiveArgumentInvalid;
    Float32* a = gather(f);
    if (!a) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = lbad_status(lbadcu_frame_decompose_host(a, f->numberOfRows, f->rowLength));
    if (e == noErr)
        for (UInt32 r = 0; r < f->numberOfRows; r++) memcpy(f->rows[r], a + (size_t)r * f->rowLength, f->rowLength * sizeof(Float32));
    free(a);
    return e;
}

OSStatus LBAudioDetectiveFrameExtractFingerprintStatus(LBAudioDetectiveFrameRef f, UInt32 inNumberOfWavelets, Boolean* outFingerprint) {  /* Frame.m:165-191 */
    if (!f || (!outFingerprint && inNumberOfWavelets)) return kLBAudioDetectiveArgumentInvalid;
    const UInt64 n = (UInt64)f->numberOfRows * f->rowLength;
    UInt32 t = inNumberOfWavelets;
    if (t > n) t = (UInt32)n;                                                          /* (upstream throws NSRangeException) */
    if (t == 0) return noErr;
    Float32* a = gather(f);
    if (!a) return kLBAudioDetectiveArgumentInvalid;
    unsigned char* set = calloc((size_t)2 * t, 1);
    OSStatus e = set ? lbad_status(lbadcu_frame_extract_host(a, (UInt32)n, t, set)) : kLBAudioDetectiveDeviceError;
    if (e == noErr)
        for (UInt32 i = 0; i < 2 * t; i++) if (set[i]) outFingerprint[i] = TRUE;       /* upstream only ever sets TRUE (Frame.m:184-189) */
    free(set); free(a);
    return e;
}

static void report(const char* what, OSStatus e) {
    fprintf(stderr, "%s: not computed (status %d): %s\n", what, (int)e, lbadcu_last_error());
}

void LBAudioDetectiveFrameDecompose(LBAudioDetectiveFrameRef f) {
    OSStatus e = LBAudioDetectiveFrameDecomposeStatus(f);
    if (e != noErr) report("LBAudioDetectiveFrameDecompose", e);
}

void LBAudioDetectiveFrameExtractFingerprint(LBAudioDetectiveFrameRef f, UInt32 inNumberOfWavelets, Boolean* outFingerprint) {
    OSStatus e = LBAudioDetectiveFrameExtractFingerprintStatus(f, inNumberOfWavelets, outFingerprint);
    if (e != noErr) report("LBAudioDetectiveFrameExtractFingerprint", e);
}

Boolean LBAudioDetectiveFrameEqualToFrame(LBAudioDetectiveFrameRef f1, LBAudioDetectiveFrameRef f2) {                                /* Frame.m:193-210 */
    if (f1->rowLength != f2->rowLength || f1->numberOfRows != f2->numberOfRows) return FALSE;
    for (UInt32 r = 0; r < f1->numberOfRows; r++)
        if (memcmp(f1->rows[r], f2->rows[r], f1->rowLength * sizeof(Float32)) != 0) return FALSE;
    return TRUE;
}
