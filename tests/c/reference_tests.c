/*
 * reference_tests.c — the reference's XCTest suite (LBAudioDetectiveTests/LBAudioDetectiveTests.m) re-stated in C against
 * the drop-in library: same API calls, same argument order (archive first, query second, range 0), with the bundled .caf
 * recordings replaced by synthetic "birds" handed over as float32 PCM (decoding is out of scope).
 *
 *   testFingerprintingWithEqualBirds      Tests.m:53-117  (suffix _eql : crop of the archive clip)       -> asserted here
 *   testFingerprintingWithDifferentBirds  (suffix _dif : another recording)                              -> printed only, as upstream
 *   testFingerprintingWithBlurredBirds    (suffix _blu1/_blu2 : crop + 1.58 % / 3.16 % noise)            -> asserted here
 *   testFingerprintVersatility            Tests.m:119-139                                                -> asserted
 *   testFingerprintComparison             Tests.m:141-155                                                -> asserted
 *   testHaarWaveletDecomposition          Tests.m:157-176  (upstream prints the 3 x 4 result)            -> asserted against the values the
 *                                                                                                           compiled reference prints
 *   (addition) the same identification from 44.1 kHz PCM, the rate of the bundled recordings, through the recording-rate entry
 *   points that stand in for ExtAudioFile's client-format conversion (LBAudioDetective.m:229)             -> asserted
 *
 * Build: cc -std=gnu11 -Iinclude tests/c/reference_tests.c -Llbaudiodetective_b200 -lLBAudioDetectiveCUDA -lm
 * Exit status 0 = all assertions hold; 77 = no CUDA device (skipped).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "LBAudioDetective.h"
#include "LBAudioDetectiveFrame.h"
#include "LBAudioDetectiveResample.h"
#include "LBAudioDetectiveSupport.h"

#define BIRDS 10
#define ARCHIVE_SAMPLES 49608   /* 9 s at 5512 Hz, like the archive clips */
#define CROP_SAMPLES 22048      /* 4 s, like the cropped clips */
#define CROP_START (8192 * 2)   /* crops start on a frame boundary of the archive clip */

static unsigned long long rng_state;
static double rnd(void) { rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(rng_state >> 40) / 16777216.0; }

/* a "bird": a few chirping partials plus a little noise, sampled at sampleRate */
static void make_bird_at(int bird, int take, Float32* out, int n, double sampleRate) {
    rng_state = 1000003ULL * (unsigned long long)(bird + 1) + 77ULL * (unsigned long long)take;
    double f[3], df[3], rate[3];
    for (int p = 0; p < 3; p++) { f[p] = 400.0 + 1400.0 * rnd(); df[p] = 100.0 + 300.0 * rnd(); rate[p] = 2.0 + 6.0 * rnd(); }
    for (int i = 0; i < n; i++) {
        double t = i / sampleRate, v = 0.0;
        for (int p = 0; p < 3; p++) v += 0.25 * sin(2.0 * M_PI * (f[p] * t + df[p] / (2.0 * M_PI * rate[p]) * sin(2.0 * M_PI * rate[p] * t)));
        out[i] = (Float32)(v + 0.05 * (rnd() - 0.5));
    }
}

static void make_bird(int bird, int take, Float32* out, int n) { make_bird_at(bird, take, out, n, 5512.0); }

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { failures++; fprintf(stderr, "FAILED %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

/* Tests.m:53-117: every archive bird against every cropped bird; returns how many birds were identified */
static int fingerprinting_with_suffix(LBAudioDetectiveRef detective, Float32 archive[BIRDS][ARCHIVE_SAMPLES], Float32 (*cropped)[CROP_SAMPLES], const char* suffix) {
    int identified = 0;
    printf("%-6s", suffix);
    for (int original = 0; original < BIRDS; original++) {
        Float32 bestMatch = 0.0f; int bestBird = -1;
        for (int sequence = 0; sequence < BIRDS; sequence++) {
            Float32 match = 0.0f;
            OSStatus error = LBAudioDetectiveComparePCM(detective, archive[original], ARCHIVE_SAMPLES, cropped[sequence], CROP_SAMPLES, 0, &match);
            CHECK(error == noErr, "LBAudioDetectiveComparePCM returned %d", (int)error);
            if (match > bestMatch) { bestMatch = match; bestBird = sequence; }          /* Tests.m:72-75 */
        }
        printf(" %d:%d(%.3f)", original, bestBird, bestMatch);
        identified += bestBird == original;
    }
    printf("  -> %d/%d identified\n", identified, BIRDS);
    return identified;
}

int main(void) {
    if (!LBAudioDetectiveSupportDeviceAvailable()) { fprintf(stderr, "no CUDA device: skipped\n"); return 77; }
    static Float32 archive[BIRDS][ARCHIVE_SAMPLES], eql[BIRDS][CROP_SAMPLES], dif[BIRDS][CROP_SAMPLES], blu1[BIRDS][CROP_SAMPLES], blu2[BIRDS][CROP_SAMPLES];
    for (int b = 0; b < BIRDS; b++) {
        make_bird(b, 0, archive[b], ARCHIVE_SAMPLES);
        memcpy(eql[b], archive[b] + CROP_START, sizeof eql[b]);
        make_bird(b, 1, dif[b], CROP_SAMPLES);
        rng_state = 4242 + b;
        for (int i = 0; i < CROP_SAMPLES; i++) { blu1[b][i] = eql[b][i] + (Float32)(0.0158 * 2.0 * (rnd() - 0.5)); blu2[b][i] = eql[b][i] + (Float32)(0.0316 * 2.0 * (rnd() - 0.5)); }
    }
    LBAudioDetectiveRef detective = LBAudioDetectiveNew();                                 /* Tests.m:42 */
    CHECK(detective != NULL, "LBAudioDetectiveNew");
    CHECK(LBAudioDetectiveGetWindowSize(detective) == 2048 && LBAudioDetectiveGetAnalysisStride(detective) == 64 &&
          LBAudioDetectiveGetNumberOfPitchSteps(detective) == 32 && LBAudioDetectiveGetSubfingerprintLength(detective) == 200, "defaults");

    CHECK(fingerprinting_with_suffix(detective, archive, eql, "_eql") == BIRDS, "a crop of the archive clip must identify its bird");
    fingerprinting_with_suffix(detective, archive, dif, "_dif");                            /* upstream asserts nothing here either */
    CHECK(fingerprinting_with_suffix(detective, archive, blu1, "_blu1") == BIRDS, "1.58 %% noise must not break identification");
    CHECK(fingerprinting_with_suffix(detective, archive, blu2, "_blu2") == BIRDS, "3.16 %% noise must not break identification");

    /* Tests.m:119-139: two detective instances produce equal fingerprints */
    for (int i = 0; i < 3; i++) {
        LBAudioDetectiveRef detective2 = LBAudioDetectiveNew();
        LBAudioDetectiveFingerprintRef fingerprint1 = NULL, fingerprint2 = NULL;
        CHECK(LBAudioDetectiveProcessPCM(detective, archive[0], ARCHIVE_SAMPLES, &fingerprint1) == noErr, "ProcessPCM");
        CHECK(LBAudioDetectiveProcessPCM(detective2, archive[0], ARCHIVE_SAMPLES, &fingerprint2) == noErr, "ProcessPCM");
        CHECK(LBAudioDetectiveFingerprintEqualToFingerprint(fingerprint1, fingerprint2), "Fingerprints should be equal");
        CHECK(LBAudioDetectiveFingerprintGetNumberOfSubfingerprints(fingerprint1) == 5 && LBAudioDetectiveFingerprintGetSubfingerprintLength(fingerprint1) == 200, "5 x 200");
        /* Tests.m:141-155: a copy equals the original; a fingerprint matches itself completely */
        LBAudioDetectiveFingerprintRef copy = LBAudioDetectiveFingerprintCopy(fingerprint1);
        CHECK(LBAudioDetectiveFingerprintEqualToFingerprint(fingerprint1, copy), "Fingerprints should be equal");
        CHECK(LBAudioDetectiveFingerprintCompareToFingerprint(fingerprint1, copy, 200) == 1.0f, "self match is 1.0");
        LBAudioDetectiveFingerprintDispose(copy); LBAudioDetectiveFingerprintDispose(fingerprint1); LBAudioDetectiveFingerprintDispose(fingerprint2);
        CHECK(LBAudioDetectiveDispose(detective2) == noErr, "Dispose");
    }
    /* the bundled recordings are 44.1 kHz files which ExtAudioFile converts while reading (m:229): here the caller hands over PCM at that
     * rate and the library converts on the GPU; archive = 9 s, query = its first 4 s.  (The comparison slides by whole subfingerprints,
     * i.e. by 8192 processing-rate samples = 65541.9 recorded samples: a crop that starts further in cannot be cut on a frame boundary of
     * the archive, and these sparse synthetic birds lose the match within a hundredth of a sample of misalignment.) */
    {
        enum { HI_ARCHIVE = 9 * 44100, HI_CROP = 4 * 44100, HI_START = 0 };
        static Float32 hi[HI_ARCHIVE];
        LBAudioDetectiveFingerprintRef fa[BIRDS], fc[BIRDS];
        CHECK(LBAudioDetectiveSetRecordingSampleRate(detective, 44100.0) == noErr, "SetRecordingSampleRate");
        CHECK(LBAudioDetectiveGetResampledLength(detective, HI_ARCHIVE) == ARCHIVE_SAMPLES, "9 s stay 9 s");
        for (int b = 0; b < BIRDS; b++) {
            make_bird_at(b, 0, hi, HI_ARCHIVE, 44100.0);
            CHECK(LBAudioDetectiveProcessRecordedPCM(detective, hi, HI_ARCHIVE, &fa[b]) == noErr, "ProcessRecordedPCM");
            CHECK(LBAudioDetectiveProcessRecordedPCM(detective, hi + HI_START, HI_CROP, &fc[b]) == noErr, "ProcessRecordedPCM");
            CHECK(LBAudioDetectiveFingerprintGetNumberOfSubfingerprints(fa[b]) == 5 && LBAudioDetectiveFingerprintGetNumberOfSubfingerprints(fc[b]) == 2, "5 and 2 subfingerprints");
        }
        int identified = 0;
        printf("%-6s", "_44k");
        for (int original = 0; original < BIRDS; original++) {
            Float32 bestMatch = 0.0f; int bestBird = -1;
            for (int sequence = 0; sequence < BIRDS; sequence++) {
                const Float32 match = LBAudioDetectiveFingerprintCompareToFingerprint(fa[original], fc[sequence], 200);
                if (match > bestMatch) { bestMatch = match; bestBird = sequence; }
            }
            printf(" %d:%d(%.3f)", original, bestBird, bestMatch);
            identified += bestBird == original;
        }
        printf("  -> %d/%d identified\n", identified, BIRDS);
        CHECK(identified == BIRDS, "a crop of a 44.1 kHz recording must identify its bird");
        for (int b = 0; b < BIRDS; b++) { LBAudioDetectiveFingerprintDispose(fa[b]); LBAudioDetectiveFingerprintDispose(fc[b]); }
    }
    /* Tests.m:157-176 through the Frame API, as written upstream; the expected values are what the compiled reference prints */
    {
        LBAudioDetectiveFrameRef frame = LBAudioDetectiveFrameNew(3);
        Float32 row1[] = {538, 940, 1940, 1794};
        Float32 row2[] = {1840, 213, 1320, 913};
        Float32 row3[] = {192, 591, 492, 1921};
        LBAudioDetectiveFrameSetRow(frame, row1, 0, 4);
        LBAudioDetectiveFrameSetRow(frame, row2, 1, 4);
        LBAudioDetectiveFrameSetRow(frame, row3, 2, 4);
        LBAudioDetectiveFrameDecompose(frame);
        static const double expected[3][4] = {{969.385559, -248.623199, 176.813522, 79.818680}, {94.509499, -211.880875, -292.860931, -37.672104},
                                              {461.302917, -235.270218, -81.445541, -291.693420}};
        int close = 1;
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 4; c++) {
                printf("%f\t", LBAudioDetectiveFrameGetValue(frame, r, c));
                close = close && fabs(LBAudioDetectiveFrameGetValue(frame, r, c) - expected[r][c]) < 5e-7 * fabs(expected[r][c]) + 1e-6;      /* six printed decimals */
            }
            printf("\n");
        }
        CHECK(close, "Haar decomposition of the 3 x 4 frame");
        Boolean signs[8] = {0};
        LBAudioDetectiveFrameExtractFingerprint(frame, 4, signs);                          /* ranks: 969.39 (+), 461.30 (+), -292.86, -291.69 */
        CHECK(signs[0] && !signs[1] && signs[2] && !signs[3] && !signs[4] && signs[5] && !signs[6] && signs[7], "signs of the four largest coefficients");
        LBAudioDetectiveFrameDispose(frame);
    }
    CHECK(LBAudioDetectiveDispose(detective) == noErr, "Dispose");                          /* Tests.m:46 */
    CHECK(LBAudioDetectiveDispose(NULL) == kLBAudioDetectiveArgumentInvalid, "Dispose(NULL)");
    printf(failures ? "%d FAILURES\n" : "all reference tests passed\n", failures);
    return failures ? 1 : 0;
}
