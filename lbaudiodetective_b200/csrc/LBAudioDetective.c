/*
 * LBAudioDetective.c — host side of the detective: configuration object, band-table construction and the
 * PCM entry points.  All signal processing is delegated to the CUDA layer (lbad_extract.cu); nothing here
 * touches a sample.  Reference: /root/reference/LBAudioDetective/LBAudioDetective.m (m:) and .h (h:).
 */
#include "lbad_host.h"
#include <pthread.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* m:20-26 (+ the one constant upstream declares but never defines, h:19) */
const OSStatus kLBAudioDetectiveArgumentInvalid = 1;
const UInt32 kLBAudioDetectiveDefaultWindowSize = 2048;
const UInt32 kLBAudioDetectiveDefaultAnalysisStride = 64;
const UInt32 kLBAudioDetectiveDefaultNumberOfPitchSteps = 32;
const UInt32 kLBAudioDetectiveDefaultNumberOfRowsPerFrame = 128;
const UInt32 kLBAudioDetectiveDefaultFingerprintComparisonRange = 0;
const UInt32 kLBAudioDetectiveDefaultSubfingerprintLength = 200;
const OSStatus kLBAudioDetectiveDeviceUnavailable = LBAD_ERR_NODEVICE;
const OSStatus kLBAudioDetectiveDeviceError = LBAD_ERR_CUDA;

/* replaces struct LBAudioDetective (m:28-44): the vDSP scratch (m:37-43) becomes a device-side plan */
struct LBAudioDetective {
    AudioStreamBasicDescription processingFormat;
    UInt32 subfingerprintLength;
    UInt32 windowSize;
    UInt32 analysisStride;
    UInt32 pitchStepCount;
    lbadcu_plan* plan;          /* built lazily; any setter invalidates it */
    UInt64 launchesBefore;      /* launches of plans that were since destroyed */
    Float64 recordingSampleRate;        /* rate of the PCM given to the ...Recorded... entry points (LBAudioDetectiveResample.h) */
    lbadcu_resampler* resampler;        /* built lazily for (recordingSampleRate, processing rate) */
    Float32* dResampled; UInt64 dResampledCapacity;     /* device scratch of ProcessRecordedPCMBatchDevice */
    int device;                 /* CUDA device of the plan, the resampler and the scratch; -1: whichever is current when they are built */
};

static void invalidate_resampler(LBAudioDetectiveRef d) {
    if (d->resampler) { d->launchesBefore += lbadcu_resampler_launches(d->resampler); lbadcu_resampler_destroy(d->resampler); d->resampler = NULL; }
}
static void invalidate_plan(LBAudioDetectiveRef d) {
    if (d->plan) { d->launchesBefore += lbadcu_plan_launches(d->plan); lbadcu_plan_destroy(d->plan); d->plan = NULL; }
}

/* m:77-90 */
LBAudioDetectiveRef LBAudioDetectiveNew(void) {
    LBAudioDetectiveRef d = calloc(1, sizeof *d);
    if (!d) return NULL;
    d->processingFormat = LBAudioDetectiveDefaultProcessingFormat();
    d->subfingerprintLength = kLBAudioDetectiveDefaultSubfingerprintLength;
    LBAudioDetectiveSetWindowSize(d, kLBAudioDetectiveDefaultWindowSize);     /* return value ignored, as upstream (m:85) */
    d->analysisStride = kLBAudioDetectiveDefaultAnalysisStride;
    d->pitchStepCount = kLBAudioDetectiveDefaultNumberOfPitchSteps;
    d->recordingSampleRate = 44100.0;
    d->device = -1;
    return d;
}

/* m:92-111 */
OSStatus LBAudioDetectiveDispose(LBAudioDetectiveRef d) {
    if (d == NULL) return kLBAudioDetectiveArgumentInvalid;
    invalidate_plan(d);
    invalidate_resampler(d);
    if (d->dResampled) lbadcu_device_free(d->dResampled);
    free(d);
    return noErr;
}

/* m:116-131 */
AudioStreamBasicDescription LBAudioDetectiveDefaultProcessingFormat(void) {
    UInt32 bytesPerSample = sizeof(Float32);
    AudioStreamBasicDescription asbd;
    memset(&asbd, 0, sizeof asbd);
    asbd.mFormatID = kAudioFormatLinearPCM;
    asbd.mFormatFlags = kAudioFormatFlagIsFloat | kAudioFormatFlagIsPacked;
    asbd.mBitsPerChannel = 8 * bytesPerSample;
    asbd.mFramesPerPacket = 1;
    asbd.mChannelsPerFrame = 1;
    asbd.mBytesPerPacket = bytesPerSample * asbd.mFramesPerPacket;
    asbd.mBytesPerFrame = bytesPerSample * asbd.mChannelsPerFrame;
    asbd.mSampleRate = 5512.0;
    return asbd;
}

/* m:133-151 */
Float64 LBAudioDetectiveGetProcessingSampleRate(LBAudioDetectiveRef d) { return d->processingFormat.mSampleRate; }
UInt32 LBAudioDetectiveGetNumberOfPitchSteps(LBAudioDetectiveRef d) { return d->pitchStepCount; }
UInt32 LBAudioDetectiveGetSubfingerprintLength(LBAudioDetectiveRef d) { return d->subfingerprintLength; }
UInt32 LBAudioDetectiveGetWindowSize(LBAudioDetectiveRef d) { return d->windowSize; }
UInt32 LBAudioDetectiveGetAnalysisStride(LBAudioDetectiveRef d) { return d->analysisStride; }

/* h:143 (declared upstream, never defined): here the rate of the PCM handed to the ...Recorded... entry points */
OSStatus LBAudioDetectiveSetRecordingSampleRate(LBAudioDetectiveRef d, Float64 inSampleRate) {
    if (!d || !(inSampleRate > 0.0)) return kLBAudioDetectiveArgumentInvalid;
    d->recordingSampleRate = inSampleRate; invalidate_resampler(d);
    return noErr;
}
Float64 LBAudioDetectiveGetRecordingSampleRate(LBAudioDetectiveRef d) { return d ? d->recordingSampleRate : 0.0; }

/* m:156-160 */
OSStatus LBAudioDetectiveSetProcessingSampleRate(LBAudioDetectiveRef d, Float64 inSampleRate) {
    d->processingFormat.mSampleRate = inSampleRate; invalidate_plan(d); invalidate_resampler(d);
    return noErr;
}
/* m:162-166 */
OSStatus LBAudioDetectiveSetNumberOfPitchSteps(LBAudioDetectiveRef d, UInt32 inNumberOfPitchSteps) {
    d->pitchStepCount = inNumberOfPitchSteps; invalidate_plan(d);
    return noErr;
}
/* m:168-172 */
OSStatus LBAudioDetectiveSetSubfingerprintLength(LBAudioDetectiveRef d, UInt32 inSubfingerprintLength) {
    d->subfingerprintLength = inSubfingerprintLength; invalidate_plan(d);
    return noErr;
}
/* m:174-195, including the inverted power-of-two check of m:183-187 (Q13) */
OSStatus LBAudioDetectiveSetWindowSize(LBAudioDetectiveRef d, UInt32 inWindowSize) {
    OSStatus error = noErr;
    invalidate_plan(d);
    d->windowSize = inWindowSize;
    UInt32 log2n = inWindowSize ? (UInt32)round(log2((double)inWindowSize)) : 0;
    UInt32 n = log2n < 32 ? (1u << log2n) : 0;
    if (n == inWindowSize) error = kLBAudioDetectiveArgumentInvalid;
    return error;
}
/* m:197-201 */
OSStatus LBAudioDetectiveSetAnalysisStride(LBAudioDetectiveRef d, UInt32 inAnalysisStride) {
    d->analysisStride = inAnalysisStride; invalidate_plan(d);
    return noErr;
}

/* m:361-371 and m:380-383, evaluated once per configuration instead of once per window */
static void band_table(LBAudioDetectiveRef d, UInt32* indices, UInt32* klow, UInt32* khigh) {
    UInt32 B = d->pitchStepCount, nFrames = d->windowSize;
    Float64 sr = d->processingFormat.mSampleRate;
    Float64 maxFreq = sr / 2.0;
    Float64 minFreq = 318.0;
    Float64 logBase = exp(log(maxFreq / minFreq) / B);
    Float64 mincoef = (Float64)d->windowSize / sr * minFreq;
    for (UInt32 j = 0; j <= B; j++) {
        UInt32 start = (UInt32)((pow(logBase, j) - 1.0) * mincoef);
        indices[j] = start + (UInt32)mincoef;
    }
    for (UInt32 i = 0; i < B; i++) {
        klow[i] = (UInt32)(((2 * indices[i]) / (sr / nFrames)) - 1);
        khigh[i] = (UInt32)(((2 * indices[i + 1]) / (sr / nFrames)) - 1);
    }
}

static int is_pow2(UInt32 v) { return v && !(v & (v - 1)); }

static OSStatus fill_geometry(LBAudioDetectiveRef d, lbadcu_geometry* g, UInt32* outIndices) {
    UInt32 N = d->windowSize, B = d->pitchStepCount, L = d->subfingerprintLength;
    Float64 sr = d->processingFormat.mSampleRate;
    if (!is_pow2(N) || N < LBAD_MIN_WINDOW || N > LBAD_MAX_WINDOW) return kLBAudioDetectiveArgumentInvalid;
    if (!is_pow2(B) || B < 4 || B > LBAD_MAX_BANDS) return kLBAudioDetectiveArgumentInvalid;
    if (d->analysisStride == 0 || !(sr > 0.0) || !isfinite(sr)) return kLBAudioDetectiveArgumentInvalid;
    if (L < 2 || (L & 1) || L > LBAD_MAX_SUBLEN || L / 2 > LBAD_ROWS_PER_FRAME * B) return kLBAudioDetectiveArgumentInvalid;
    /* guard the double -> UInt32 conversions of m:369-370, m:382-383 (undefined when negative or huge) */
    Float64 mincoef = (Float64)N / sr * 318.0;
    if (!(mincoef >= 1.0) || mincoef > 1e6 || sr / 2.0 <= 318.0) return kLBAudioDetectiveArgumentInvalid;
    if (((2.0 * (UInt32)mincoef) / (sr / N)) - 1.0 < 0.0) return kLBAudioDetectiveArgumentInvalid;
    UInt32 indices[LBAD_MAX_BANDS + 1];
    memset(g, 0, sizeof *g);
    band_table(d, indices, g->klow, g->khigh);
    for (UInt32 i = 0; i < B; i++) {
        if (g->klow[i] > g->khigh[i] || g->khigh[i] > N / 2) return kLBAudioDetectiveArgumentInvalid;   /* Q15 */
        g->divisor[i] = (Float32)(indices[i + 1] - indices[i]);                                          /* m:404 */
    }
    UInt32 width = (UInt32)(N / 2.0);                                                                    /* m:373 */
    g->pos_scale = (Float32)(width / 2);                                                                 /* m:391 */
    g->window = N; g->stride = d->analysisStride; g->bands = B; g->sublen = L;
    if (outIndices) memcpy(outIndices, indices, (B + 1) * sizeof(UInt32));
    return noErr;
}

static OSStatus ensure_plan(LBAudioDetectiveRef d) {
    if (d->plan) return noErr;
    lbadcu_geometry g;
    OSStatus e = fill_geometry(d, &g, NULL);
    if (e != noErr) return e;
    int prev;
    e = lbad_status(lbadcu_push_device(d->device, &prev));
    if (e != noErr) return e;
    e = lbad_status(lbadcu_plan_create(&g, &d->plan));
    lbadcu_pop_device(prev);
    return e;
}

/* Additions: which CUDA device a detective computes on.  The default (-1) is the device that is current when the first computing call
 * builds the device plan; a caller that does not use the CUDA runtime itself picks one here.  Changing it drops the plan. */
OSStatus LBAudioDetectiveSetDevice(LBAudioDetectiveRef d, int inDevice) {
    if (!d || inDevice < -1) return kLBAudioDetectiveArgumentInvalid;
    if (inDevice >= 0) {
        if (lbadcu_device_available() != LBAD_OK) return kLBAudioDetectiveDeviceUnavailable;
        if (inDevice >= lbadcu_device_count()) return kLBAudioDetectiveArgumentInvalid;
    }
    if (inDevice != d->device) {
        invalidate_plan(d); invalidate_resampler(d);
        if (d->dResampled) { lbadcu_device_free(d->dResampled); d->dResampled = NULL; d->dResampledCapacity = 0; }
        d->device = inDevice;
    }
    return noErr;
}
int LBAudioDetectiveGetDevice(LBAudioDetectiveRef d) {
    if (!d) return -1;
    return d->plan ? lbadcu_plan_device(d->plan) : d->device;
}

OSStatus LBAudioDetectiveCheckConfiguration(LBAudioDetectiveRef d) {
    if (!d) return kLBAudioDetectiveArgumentInvalid;
    lbadcu_geometry g;
    return fill_geometry(d, &g, NULL);
}

UInt64 LBAudioDetectiveGetNumberOfSubfingerprintsForLength(LBAudioDetectiveRef d, UInt64 n) {
    if (!d || n < d->windowSize || d->analysisStride == 0) return 0;
    UInt64 imageWidth = (n - d->windowSize) / d->analysisStride;                 /* m:250 */
    return imageWidth / kLBAudioDetectiveDefaultNumberOfRowsPerFrame;            /* m:255 */
}

OSStatus LBAudioDetectiveGetBandTable(LBAudioDetectiveRef d, UInt32* outIndices, UInt32* outLow, UInt32* outHigh) {
    if (!d) return kLBAudioDetectiveArgumentInvalid;
    lbadcu_geometry g; UInt32 indices[LBAD_MAX_BANDS + 1];
    OSStatus e = fill_geometry(d, &g, indices);
    if (e != noErr) return e;
    if (outIndices) memcpy(outIndices, indices, (d->pitchStepCount + 1) * sizeof(UInt32));
    if (outLow) memcpy(outLow, g.klow, d->pitchStepCount * sizeof(UInt32));
    if (outHigh) memcpy(outHigh, g.khigh, d->pitchStepCount * sizeof(UInt32));
    return noErr;
}

/* replaces m:208-308 */
OSStatus LBAudioDetectiveProcessPCM(LBAudioDetectiveRef d, const Float32* inSamples, UInt64 inNumberFrames, LBAudioDetectiveFingerprintRef* outFingerprint) {
    if (!d || !outFingerprint) return kLBAudioDetectiveArgumentInvalid;
    *outFingerprint = NULL;
    if (!inSamples) return kLBAudioDetectiveArgumentInvalid;                     /* m:211-214 */
    OSStatus e = ensure_plan(d);
    if (e != noErr) return e;
    LBAudioDetectiveFingerprintRef fp = LBAudioDetectiveFingerprintNew(0);       /* m:297 */
    UInt32 L = d->subfingerprintLength;
    LBAudioDetectiveFingerprintSetSubfingerprintLength(fp, &L);                  /* m:326-327 */
    *outFingerprint = fp;
    if (inNumberFrames < d->windowSize) return kLBAudioDetectiveArgumentInvalid; /* upstream underflows at m:250 */
    UInt64 count = LBAudioDetectiveGetNumberOfSubfingerprintsForLength(d, inNumberFrames);
    if (count == 0) return noErr;
    if (count > 0x7fffffffu) return kLBAudioDetectiveArgumentInvalid;
    UInt32 W = lbad_words_per_plane(L);
    UInt32* words = malloc((size_t)count * 2 * W * sizeof(UInt32));
    if (!words) return kLBAudioDetectiveArgumentInvalid;
    e = lbad_status(lbadcu_extract_host(d->plan, inSamples, 1, inNumberFrames, inNumberFrames, words, NULL, NULL, 0));
    if (e == noErr) e = lbad_fingerprint_append_packed(fp, words, (UInt32)count);   /* m:328 */
    free(words);
    return e;
}

/* replaces m:442-464 */
OSStatus LBAudioDetectiveComparePCM(LBAudioDetectiveRef d, const Float32* s1, UInt64 n1, const Float32* s2, UInt64 n2, UInt32 inComparisonRange, Float32* outMatch) {
    if (!d) return kLBAudioDetectiveArgumentInvalid;
    if (inComparisonRange == 0) inComparisonRange = d->subfingerprintLength;     /* m:443-445 */
    /* the usual case — both clips yield subfingerprints — runs as one device pipeline: upload, fingerprint, compare, four bytes back.
     * Same arithmetic as the three steps below (the same kernels on the same words), without their host round trips. */
    if (s1 && s2 && ensure_plan(d) == noErr && LBAudioDetectiveGetNumberOfSubfingerprintsForLength(d, n1) >= 1 &&
        LBAudioDetectiveGetNumberOfSubfingerprintsForLength(d, n2) >= 1 && LBAudioDetectiveGetNumberOfSubfingerprintsForLength(d, n1) +
        LBAudioDetectiveGetNumberOfSubfingerprintsForLength(d, n2) <= 0x7fffffffu) {
        Float32 match = 0.0f;
        OSStatus e = lbad_status(lbadcu_compare_pcm_host(d->plan, s1, n1, s2, n2, lbad_pairs_for_range(inComparisonRange, d->subfingerprintLength), &match));
        if (e == noErr && outMatch) *outMatch = match;                           /* m:456-458 */
        return e;
    }
    LBAudioDetectiveFingerprintRef fp1 = NULL, fp2 = NULL;
    OSStatus error = LBAudioDetectiveProcessPCM(d, s1, n1, &fp1);                /* m:449 */
    if (error == noErr) error = LBAudioDetectiveProcessPCM(d, s2, n2, &fp2);     /* m:453 */
    if (error == noErr && outMatch) *outMatch = LBAudioDetectiveFingerprintCompareToFingerprint(fp1, fp2, inComparisonRange);   /* m:456-458 */
    LBAudioDetectiveFingerprintDispose(fp1);                                     /* m:460-461 */
    LBAudioDetectiveFingerprintDispose(fp2);
    return error;
}

OSStatus LBAudioDetectiveProcessPCMBatch(LBAudioDetectiveRef d, const Float32* inSamples, UInt32 nClips, UInt64 framesPerClip, UInt64 clipStride, UInt32* outWords) {
    if (!d || !inSamples || !outWords || nClips == 0 || clipStride < framesPerClip) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = ensure_plan(d);
    if (e != noErr) return e;
    if (framesPerClip < d->windowSize) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_extract_host(d->plan, inSamples, nClips, framesPerClip, clipStride, outWords, NULL, NULL, 0));
}

/* LBAudioDetectiveProcessPCMBatchSharded: every detective's host thread runs the chunked upload / kernel / download pipeline of
 * LBAudioDetectiveProcessPCMBatch on the SAME batch, taking its chunks (about 190 MB of PCM) from one shared cursor as its buffers
 * drain.  The GPUs of one box do not reach the host at the same speed — on the 8-GPU B200 box of this pool four links deliver 23 GB/s
 * and four 35 GB/s when all copy at once — so equal shares would wait for the slowest link; this way the shares follow the links. */
struct shard_job { LBAudioDetectiveRef d; const Float32* samples; UInt32 nClips; UInt64 framesPerClip, clipStride; UInt32* words; uint64_t* cursor; OSStatus status; char message[256]; };
static void* shard_main(void* arg) {
    struct shard_job* j = arg;
    j->status = lbad_status(lbadcu_extract_host_shared(j->d->plan, j->samples, j->nClips, j->framesPerClip, j->clipStride, j->words, j->cursor));
    j->message[0] = 0;
    if (j->status != noErr) snprintf(j->message, sizeof j->message, "%s", lbadcu_last_error());      /* the error text lives in this thread */
    return NULL;
}

OSStatus LBAudioDetectiveProcessPCMBatchSharded(const LBAudioDetectiveRef* dets, UInt32 nDets, const Float32* inSamples, UInt32 nClips, UInt64 framesPerClip, UInt64 clipStride, UInt32* outWords) {
    if (!dets || nDets == 0 || nDets > 64 || !inSamples || !outWords || nClips == 0 || clipStride < framesPerClip) return kLBAudioDetectiveArgumentInvalid;
    for (UInt32 i = 0; i < nDets; i++) {
        if (!dets[i]) return kLBAudioDetectiveArgumentInvalid;
        if (dets[i]->windowSize != dets[0]->windowSize || dets[i]->analysisStride != dets[0]->analysisStride || dets[i]->pitchStepCount != dets[0]->pitchStepCount ||
            dets[i]->subfingerprintLength != dets[0]->subfingerprintLength || dets[i]->processingFormat.mSampleRate != dets[0]->processingFormat.mSampleRate)
            return kLBAudioDetectiveArgumentInvalid;
        for (UInt32 k = 0; k < i; k++) if (dets[k] == dets[i]) return kLBAudioDetectiveArgumentInvalid;      /* a detective is not re-entrant (one per thread) */
    }
    if (framesPerClip < dets[0]->windowSize) return kLBAudioDetectiveArgumentInvalid;
    for (UInt32 i = 0; i < nDets; i++) { OSStatus e = ensure_plan(dets[i]); if (e != noErr) return e; }
    uint64_t cursor[2] = {0, 0};                                                  /* clips handed out so far; clips per chunk (fixed by the first pipeline to start) */
    struct shard_job jobs[64]; pthread_t threads[64]; int started[64];
    for (UInt32 i = 0; i < nDets; i++) {
        jobs[i] = (struct shard_job){dets[i], inSamples, nClips, framesPerClip, clipStride, outWords, cursor, noErr, {0}};
        started[i] = pthread_create(&threads[i], NULL, shard_main, &jobs[i]) == 0;
        if (!started[i]) shard_main(&jobs[i]);                                    /* no thread to be had: this detective works here (and may take everything) */
    }
    OSStatus e = noErr;
    for (UInt32 i = 0; i < nDets; i++) {
        if (started[i]) pthread_join(threads[i], NULL);
        if (e == noErr && jobs[i].status != noErr) { e = jobs[i].status; lbadcu_set_last_error(jobs[i].message); }
    }
    return e;
}

OSStatus LBAudioDetectiveProcessPCMBatchInt16(LBAudioDetectiveRef d, const SInt16* inSamples, UInt32 nClips, UInt64 framesPerClip, UInt64 clipStride, UInt32* outWords) {
    if (!d || !inSamples || !outWords || nClips == 0 || clipStride < framesPerClip) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = ensure_plan(d);
    if (e != noErr) return e;
    if (framesPerClip < d->windowSize) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_extract_host_i16(d->plan, inSamples, nClips, framesPerClip, clipStride, outWords));
}

OSStatus LBAudioDetectiveProcessPCMBatchDevice(LBAudioDetectiveRef d, const Float32* dSamples, UInt32 nClips, UInt64 framesPerClip, UInt64 clipStride, UInt32* dWords, void* stream) {
    if (!d || !dSamples || !dWords || nClips == 0 || clipStride < framesPerClip) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = ensure_plan(d);
    if (e != noErr) return e;
    if (framesPerClip < d->windowSize) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_extract_device(d->plan, dSamples, nClips, framesPerClip, clipStride, dWords, NULL, NULL, 0, stream));
}

OSStatus LBAudioDetectiveProcessPCMStages(LBAudioDetectiveRef d, const Float32* inSamples, UInt64 n, Float32* outImages, Float32* outHaar, Boolean* outBooleans, Boolean useFused) {
    if (!d || !inSamples) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = ensure_plan(d);
    if (e != noErr) return e;
    if (n < d->windowSize) return kLBAudioDetectiveArgumentInvalid;
    if (useFused && !lbadcu_plan_fused_supported(d->plan)) return kLBAudioDetectiveArgumentInvalid;
    UInt64 count = LBAudioDetectiveGetNumberOfSubfingerprintsForLength(d, n);
    if (count == 0) return noErr;
    UInt32 L = d->subfingerprintLength, W = lbad_words_per_plane(L);
    UInt32* words = malloc((size_t)count * 2 * W * sizeof(UInt32));
    if (!words) return kLBAudioDetectiveArgumentInvalid;
    e = lbad_status(lbadcu_extract_host(d->plan, inSamples, 1, n, n, words, outImages, outHaar, useFused ? 1 : 2));
    if (e == noErr && outBooleans) for (UInt64 i = 0; i < count; i++) lbad_unpack_words(words + i * 2 * W, L, W, outBooleans + i * L);
    free(words);
    return e;
}

OSStatus LBAudioDetectiveProcessPCMBatchStages(LBAudioDetectiveRef d, const Float32* inSamples, UInt32 nClips, UInt64 framesPerClip, UInt64 clipStride,
                                               UInt32* outWords, Float32* outImages, Float32* outHaar, Boolean useFused) {
    if (!d || !inSamples || !outWords || nClips == 0 || clipStride < framesPerClip) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = ensure_plan(d);
    if (e != noErr) return e;
    if (framesPerClip < d->windowSize) return kLBAudioDetectiveArgumentInvalid;
    if (useFused && !lbadcu_plan_fused_supported(d->plan)) return kLBAudioDetectiveArgumentInvalid;
    return lbad_status(lbadcu_extract_host(d->plan, inSamples, nClips, framesPerClip, clipStride, outWords, outImages, outHaar, useFused ? 1 : 2));
}

OSStatus LBAudioDetectiveTransformImages(LBAudioDetectiveRef d, const Float32* inImages, UInt32 inCount, Float32* outHaar, Boolean* outBooleans) {
    if (!d || !inImages || inCount == 0) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = ensure_plan(d);
    if (e != noErr) return e;
    UInt32 L = d->subfingerprintLength, W = lbad_words_per_plane(L);
    UInt32* words = malloc((size_t)inCount * 2 * W * sizeof(UInt32));
    if (!words) return kLBAudioDetectiveArgumentInvalid;
    e = lbad_status(lbadcu_transform_images_host(d->plan, inImages, inCount, outHaar, words));
    if (e == noErr && outBooleans) for (UInt32 i = 0; i < inCount; i++) lbad_unpack_words(words + (size_t)i * 2 * W, L, W, outBooleans + (size_t)i * L);
    free(words);
    return e;
}

/* ---- recording-rate front end (include/LBAudioDetectiveResample.h; replaces the ExtAudioFile client-format conversion, m:229 / m:275) ---- */

static OSStatus ensure_resampler(LBAudioDetectiveRef d) {
    if (d->resampler) return noErr;
    lbadcu_resample_design des;
    if (lbad_resample_design_create(d->recordingSampleRate, d->processingFormat.mSampleRate, &des) != LBAD_OK) return kLBAudioDetectiveArgumentInvalid;
    int prev;                                                                      /* on the plan's device (or the one chosen with LBAudioDetectiveSetDevice) */
    OSStatus e = lbad_status(lbadcu_push_device(d->plan ? lbadcu_plan_device(d->plan) : d->device, &prev));
    if (e == noErr) { e = lbad_status(lbadcu_resampler_create(&des, &d->resampler)); lbadcu_pop_device(prev); }
    lbad_resample_design_free(&des);
    return e;
}

UInt64 LBAudioDetectiveGetResampledLength(LBAudioDetectiveRef d, UInt64 inNumberFrames) {
    if (!d) return 0;
    return lbad_resample_out_len(d->recordingSampleRate, d->processingFormat.mSampleRate, inNumberFrames);
}

OSStatus LBAudioDetectiveResamplePCM(LBAudioDetectiveRef d, const Float32* inSamples, UInt64 inNumberFrames, Float32* outSamples) {
    if (!d || !inSamples || !outSamples) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = ensure_resampler(d);
    if (e != noErr) return e;
    return lbad_status(lbadcu_resample_host(d->resampler, inSamples, inNumberFrames, outSamples, LBAudioDetectiveGetResampledLength(d, inNumberFrames)));
}

OSStatus LBAudioDetectiveProcessRecordedPCMBatchDevice(LBAudioDetectiveRef d, const Float32* dSamples, UInt32 nClips, UInt64 framesPerClip, UInt64 clipStride, UInt32* dWords, void* stream) {
    if (!d || !dSamples || !dWords || nClips == 0 || clipStride < framesPerClip) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = ensure_plan(d);
    if (e == noErr) e = ensure_resampler(d);
    if (e != noErr) return e;
    const UInt64 outLen = LBAudioDetectiveGetResampledLength(d, framesPerClip);
    if (outLen < d->windowSize) return kLBAudioDetectiveArgumentInvalid;
    const UInt64 outStride = (outLen + 3) & ~(UInt64)3;                            /* keeps every clip 16-byte aligned for the bulk copies of the FFT kernel */
    if (d->dResampledCapacity < outStride * nClips) {
        if (d->dResampled) lbadcu_device_free(d->dResampled);
        d->dResampled = NULL; d->dResampledCapacity = 0;
        int prev;
        e = lbad_status(lbadcu_push_device(lbadcu_plan_device(d->plan), &prev));
        if (e == noErr) { e = lbad_status(lbadcu_device_alloc_floats(outStride * nClips, &d->dResampled)); lbadcu_pop_device(prev); }
        if (e != noErr) return e;
        d->dResampledCapacity = outStride * nClips;
    }
    if (!stream) stream = lbadcu_plan_stream(d->plan);                            /* both kernels on one stream: the extraction follows the conversion */
    e = lbad_status(lbadcu_resample_device(d->resampler, dSamples, nClips, framesPerClip, clipStride, d->dResampled, outLen, outStride, stream));
    if (e != noErr) return e;
    return lbad_status(lbadcu_extract_device(d->plan, d->dResampled, nClips, outLen, outStride, dWords, NULL, NULL, 0, stream));
}

OSStatus LBAudioDetectiveProcessRecordedPCM(LBAudioDetectiveRef d, const Float32* inSamples, UInt64 inNumberFrames, LBAudioDetectiveFingerprintRef* outFingerprint) {
    if (!d || !outFingerprint) return kLBAudioDetectiveArgumentInvalid;
    *outFingerprint = NULL;
    if (!inSamples) return kLBAudioDetectiveArgumentInvalid;
    OSStatus e = ensure_plan(d);
    if (e == noErr) e = ensure_resampler(d);
    if (e != noErr) return e;
    LBAudioDetectiveFingerprintRef fp = LBAudioDetectiveFingerprintNew(0);
    UInt32 L = d->subfingerprintLength;
    LBAudioDetectiveFingerprintSetSubfingerprintLength(fp, &L);
    *outFingerprint = fp;
    const UInt64 outLen = LBAudioDetectiveGetResampledLength(d, inNumberFrames);
    if (outLen < d->windowSize) return kLBAudioDetectiveArgumentInvalid;
    const UInt64 count = LBAudioDetectiveGetNumberOfSubfingerprintsForLength(d, outLen);
    if (count == 0) return noErr;
    if (count > 0x7fffffffu) return kLBAudioDetectiveArgumentInvalid;
    const UInt32 W = lbad_words_per_plane(L);
    UInt32* words = malloc((size_t)count * 2 * W * sizeof(UInt32));
    if (!words) return kLBAudioDetectiveArgumentInvalid;
    e = lbad_status(lbadcu_process_recorded_host(d->plan, d->resampler, inSamples, inNumberFrames, outLen, words, (size_t)count * 2 * W));
    if (e == noErr) e = lbad_fingerprint_append_packed(fp, words, (UInt32)count);
    free(words);
    return e;
}

UInt64 LBAudioDetectiveGetKernelLaunchCount(LBAudioDetectiveRef d) {
    if (!d) return 0;
    return d->launchesBefore + (d->plan ? lbadcu_plan_launches(d->plan) : 0) + (d->resampler ? lbadcu_resampler_launches(d->resampler) : 0);
}

UInt32 LBAudioDetectiveGetKernelTiming(LBAudioDetectiveRef d, Boolean inEnable, Boolean inReset, Float64* outTotalMilliseconds) {
    if (outTotalMilliseconds) *outTotalMilliseconds = 0.0;
    if (!d || ensure_plan(d) != noErr) return 0;
    return lbadcu_plan_timing(d->plan, 0, inEnable, inReset, outTotalMilliseconds);
}

UInt32 LBAudioDetectiveGetTransformKernelTiming(LBAudioDetectiveRef d, Boolean inEnable, Boolean inReset, Float64* outTotalMilliseconds) {
    if (outTotalMilliseconds) *outTotalMilliseconds = 0.0;
    if (!d || ensure_plan(d) != noErr) return 0;
    return lbadcu_plan_timing(d->plan, 1, inEnable, inReset, outTotalMilliseconds);
}

/* ---- streaming extraction (addition): append PCM, a subfingerprint is emitted for every completed frame ---- */

struct LBAudioDetectiveStream {
    LBAudioDetectiveRef detective;              /* borrowed: configuration + device plan */
    LBAudioDetectiveFingerprintRef fingerprint; /* grows as frames complete */
    Float32* pending;                           /* samples from the start of the first frame not yet emitted */
    UInt64 pendingCount, pendingCapacity;
    UInt64 totalFrames;                         /* samples appended so far */
    UInt64 emitted;                             /* subfingerprints emitted so far */
    UInt32 window, stride, sublen, bands;       /* configuration frozen at creation: a change on the borrowed detective fails the next append */
    Float64 sampleRate;
};

LBAudioDetectiveStreamRef LBAudioDetectiveStreamNew(LBAudioDetectiveRef d) {
    if (!d || LBAudioDetectiveCheckConfiguration(d) != noErr) return NULL;
    LBAudioDetectiveStreamRef s = calloc(1, sizeof *s);
    if (!s) return NULL;
    s->detective = d; s->window = d->windowSize; s->stride = d->analysisStride; s->sublen = d->subfingerprintLength;
    s->bands = d->pitchStepCount; s->sampleRate = d->processingFormat.mSampleRate;
    s->fingerprint = LBAudioDetectiveFingerprintNew(0);
    UInt32 L = s->sublen;
    LBAudioDetectiveFingerprintSetSubfingerprintLength(s->fingerprint, &L);
    return s;
}

OSStatus LBAudioDetectiveStreamDispose(LBAudioDetectiveStreamRef s) {
    if (!s) return kLBAudioDetectiveArgumentInvalid;
    LBAudioDetectiveFingerprintDispose(s->fingerprint);
    free(s->pending); free(s);
    return noErr;
}

OSStatus LBAudioDetectiveStreamAppend(LBAudioDetectiveStreamRef s, const Float32* inSamples, UInt64 inNumberFrames) {
    if (!s || (!inSamples && inNumberFrames)) return kLBAudioDetectiveArgumentInvalid;
    LBAudioDetectiveRef d = s->detective;
    /* subfingerprints of one fingerprint must come from one band table and one geometry */
    if (d->windowSize != s->window || d->analysisStride != s->stride || d->subfingerprintLength != s->sublen || d->pitchStepCount != s->bands ||
        d->processingFormat.mSampleRate != s->sampleRate) return kLBAudioDetectiveArgumentInvalid;
    if (s->pendingCount + inNumberFrames > s->pendingCapacity) {
        UInt64 cap = s->pendingCapacity ? s->pendingCapacity * 2 : 65536;
        while (cap < s->pendingCount + inNumberFrames) cap *= 2;
        Float32* p = realloc(s->pending, cap * sizeof(Float32));
        if (!p) return kLBAudioDetectiveArgumentInvalid;
        s->pending = p; s->pendingCapacity = cap;
    }
    memcpy(s->pending + s->pendingCount, inSamples, inNumberFrames * sizeof(Float32));
    s->pendingCount += inNumberFrames; s->totalFrames += inNumberFrames;
    /* frames the one-shot path would produce for everything appended so far (m:250-255), minus those already emitted;
     * the pending buffer starts at sample 128*stride*emitted, so processing it as one clip yields exactly the new ones */
    UInt64 available = LBAudioDetectiveGetNumberOfSubfingerprintsForLength(d, s->totalFrames);
    if (available <= s->emitted) return noErr;
    UInt64 fresh = available - s->emitted;
    LBAudioDetectiveFingerprintRef part = NULL;
    OSStatus e = LBAudioDetectiveProcessPCM(d, s->pending, s->pendingCount, &part);
    if (e == noErr && part->subfingerprintCount != fresh) e = kLBAudioDetectiveDeviceError;
    if (e == noErr) e = lbad_fingerprint_append_packed(s->fingerprint, part->words, part->subfingerprintCount);
    LBAudioDetectiveFingerprintDispose(part);
    if (e != noErr) return e;
    UInt64 consumed = fresh * kLBAudioDetectiveDefaultNumberOfRowsPerFrame * s->stride;
    memmove(s->pending, s->pending + consumed, (s->pendingCount - consumed) * sizeof(Float32));
    s->pendingCount -= consumed; s->emitted = available;
    return noErr;
}

LBAudioDetectiveFingerprintRef LBAudioDetectiveStreamGetFingerprint(LBAudioDetectiveStreamRef s) { return s ? s->fingerprint : NULL; }
UInt64 LBAudioDetectiveStreamGetNumberOfPendingFrames(LBAudioDetectiveStreamRef s) { return s ? s->pendingCount : 0; }
