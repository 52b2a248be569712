/*
 * sharded_extract.c — a plain-C caller of the drop-in library fingerprinting one batch of clips on several GPUs through the library
 * alone: one detective per device (LBAudioDetectiveSetDevice), one call (LBAudioDetectiveProcessPCMBatchSharded).  The reference's
 * callers loop over files with LBAudioDetectiveProcessAudioURL (LBAudioDetectiveTests.m:53-60); clips are independent, so the
 * multi-GPU form of that loop shards by clip and exchanges nothing.
 *
 * The detectives go to the visible CUDA devices round-robin (three detectives share the one device of a single-GPU box).  Checks:
 *   - the sharded result equals what ONE detective returns for the same batch, word for word, for a clip count that does not divide evenly,
 *   - every fingerprint equals LBAudioDetectiveProcessPCM of its clip,
 *   - detectives that are configured differently are refused.
 *
 * Build: cc -std=gnu11 -Iinclude tests/c/sharded_extract.c -Llbaudiodetective_b200 -lLBAudioDetectiveCUDA -lm
 * Exit status 0 = all checks hold; 77 = no CUDA device (skipped).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "LBAudioDetective.h"
#include "LBAudioDetectiveSupport.h"

#define DETECTIVES 3
#define CLIPS 7
#define CLIP_SAMPLES 55120      /* 10 s: 6 subfingerprints */
#define SUBFPS 6
#define W2 8

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { failures++; fprintf(stderr, "FAILED %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

int main(void) {
    if (!LBAudioDetectiveSupportDeviceAvailable()) { fprintf(stderr, "no CUDA device: skipped\n"); return 77; }
    const int n_dev = LBAudioDetectiveSupportDeviceCount();
    static Float32 pcm[CLIPS][CLIP_SAMPLES];
    unsigned long long s = 12345;
    for (int c = 0; c < CLIPS; c++)
        for (int i = 0; i < CLIP_SAMPLES; i++) {
            s = s * 6364136223846793005ULL + 1442695040888963407ULL;
            const double t = i / 5512.0, f0 = 300.0 + 90.0 * c, f1 = 1500.0 + 60.0 * c;
            pcm[c][i] = (Float32)(0.5 * sin(2.0 * M_PI * (f0 * t + (f1 - f0) * t * t / 20.0)) + 0.2 * sin(2.0 * M_PI * (500.0 + 130.0 * c) * t) + 0.1 * ((double)(s >> 40) / 16777216.0 - 0.5));
        }
    LBAudioDetectiveRef dets[DETECTIVES];
    for (int i = 0; i < DETECTIVES; i++) {
        dets[i] = LBAudioDetectiveNew();
        CHECK(LBAudioDetectiveGetDevice(dets[i]) == -1, "a new detective has no device yet");
        CHECK(LBAudioDetectiveSetDevice(dets[i], i % n_dev) == noErr, "SetDevice(%d)", i % n_dev);
        CHECK(LBAudioDetectiveGetDevice(dets[i]) == i % n_dev, "GetDevice");
    }
    CHECK(LBAudioDetectiveSetDevice(dets[0], n_dev) == kLBAudioDetectiveArgumentInvalid, "a device that does not exist is refused");

    static UInt32 sharded[CLIPS][SUBFPS][W2], single[CLIPS][SUBFPS][W2];
    CHECK(LBAudioDetectiveProcessPCMBatchSharded(dets, DETECTIVES, &pcm[0][0], CLIPS, CLIP_SAMPLES, CLIP_SAMPLES, &sharded[0][0][0]) == noErr, "ProcessPCMBatchSharded");
    LBAudioDetectiveRef one = LBAudioDetectiveNew();
    CHECK(LBAudioDetectiveProcessPCMBatch(one, &pcm[0][0], CLIPS, CLIP_SAMPLES, CLIP_SAMPLES, &single[0][0][0]) == noErr, "ProcessPCMBatch");
    CHECK(memcmp(sharded, single, sizeof sharded) == 0, "the sharded batch must equal the one-detective batch");
    for (int c = 0; c < CLIPS; c++) {
        LBAudioDetectiveFingerprintRef fp = NULL;
        CHECK(LBAudioDetectiveProcessPCM(one, pcm[c], CLIP_SAMPLES, &fp) == noErr && LBAudioDetectiveFingerprintGetNumberOfSubfingerprints(fp) == SUBFPS, "ProcessPCM");
        for (UInt32 j = 0; j < SUBFPS; j++) {
            UInt32 w[W2];
            CHECK(LBAudioDetectiveFingerprintGetPackedSubfingerprintAtIndex(fp, j, w) == W2 && memcmp(w, sharded[c][j], sizeof w) == 0, "clip %d subfingerprint %u", c, (unsigned)j);
        }
        LBAudioDetectiveFingerprintDispose(fp);
    }
    /* a share of zero clips (more detectives than clips) and mismatching configurations */
    static UInt32 two[2][SUBFPS][W2];
    CHECK(LBAudioDetectiveProcessPCMBatchSharded(dets, DETECTIVES, &pcm[0][0], 2, CLIP_SAMPLES, CLIP_SAMPLES, &two[0][0][0]) == noErr && memcmp(two, single, sizeof two) == 0, "two clips over three detectives");
    LBAudioDetectiveSetSubfingerprintLength(dets[1], 100);
    CHECK(LBAudioDetectiveProcessPCMBatchSharded(dets, DETECTIVES, &pcm[0][0], CLIPS, CLIP_SAMPLES, CLIP_SAMPLES, &sharded[0][0][0]) == kLBAudioDetectiveArgumentInvalid, "differently configured detectives are refused");
    LBAudioDetectiveRef twice[2] = {dets[0], dets[0]};
    CHECK(LBAudioDetectiveProcessPCMBatchSharded(twice, 2, &pcm[0][0], CLIPS, CLIP_SAMPLES, CLIP_SAMPLES, &sharded[0][0][0]) == kLBAudioDetectiveArgumentInvalid, "one detective cannot take two shares");
    for (int i = 0; i < DETECTIVES; i++) LBAudioDetectiveDispose(dets[i]);
    LBAudioDetectiveDispose(one);
    printf("%d clips over %d detectives on %d device(s): sharded batch identical to one detective's\n", CLIPS, DETECTIVES, n_dev);
    printf(failures ? "%d FAILURES\n" : "all checks passed\n", failures);
    return failures ? 1 : 0;
}
