/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Flat, ctypes-friendly entry points around the COMPILED REFERENCE (oracle/_ref/libLBAudioDetectiveRef.so):
 * every function here only marshals buffers and then calls the reference's own functions
 * (LBAudioDetective.m / LBAudioDetectiveFrame.m / LBAudioDetectiveFingerprint.m, built by oracle/Makefile).
 * Used to (1) pin the C restatement in oracle/lbad_oracle.c, (2) generate tests/golden/, (3) time the
 * reference on host cores (bench.py --impl reference, cpu_baseline.kind = "reference").
 */
#include <Foundation/Foundation.h>
#include <AudioToolbox/AudioToolbox.h>
#include <Accelerate/Accelerate.h>
#include <pthread.h>
#include <time.h>
#include "LBAudioDetective.h"
#include "LBAudioDetectiveFrame.h"
#include "LBAudioDetectiveFingerprint.h"

/* non-static internals of the reference (LBAudioDetective.m:46-47) */
void LBAudioDetectiveSynthesizeFingerprint(LBAudioDetectiveRef, LBAudioDetectiveFrameRef*, UInt64, LBAudioDetectiveFingerprintRef*);
OSStatus LBAudioDetectiveComputeFrequencies(LBAudioDetectiveRef, void*, UInt32, AudioStreamBasicDescription, UInt32, Float32*);
extern const UInt32 kLBAudioDetectiveDefaultNumberOfRowsPerFrame;

struct LBADShimURL { const float* samples; SInt64 count; };   /* must match oracle/shim/shim.c */

typedef struct { UInt32 window, stride, bands, sublen; Float64 sample_rate; } lbad_ref_cfg;

static LBAudioDetectiveRef make_detective(const lbad_ref_cfg* c) {
    LBAudioDetectiveRef d = LBAudioDetectiveNew();
    if (c) {
        if (c->sample_rate > 0) LBAudioDetectiveSetProcessingSampleRate(d, c->sample_rate);
        if (c->window) LBAudioDetectiveSetWindowSize(d, c->window);
        if (c->stride) LBAudioDetectiveSetAnalysisStride(d, c->stride);
        if (c->bands) LBAudioDetectiveSetNumberOfPitchSteps(d, c->bands);
        if (c->sublen) LBAudioDetectiveSetSubfingerprintLength(d, c->sublen);
    }
    return d;
}

static UInt32 dump_fingerprint(LBAudioDetectiveFingerprintRef fp, Boolean* out, UInt32 max_subfps, UInt32* out_len) {
    UInt32 n = LBAudioDetectiveFingerprintGetNumberOfSubfingerprints(fp);
    UInt32 L = LBAudioDetectiveFingerprintGetSubfingerprintLength(fp);
    if (out_len) *out_len = L;
    for (UInt32 i = 0; i < n && i < max_subfps; i++) LBAudioDetectiveFingerprintGetSubfingerprintAtIndex(fp, i, out + (size_t)i * L);
    return n;
}

/* ProcessAudioURL exactly as written (LBAudioDetective.m:208-308) on a memory-backed "file". */
int lbad_ref_process_pcm(const lbad_ref_cfg* cfg, const float* pcm, SInt64 n, Boolean* out_bits, UInt32 max_subfps,
                         UInt32* out_count, UInt32* out_len) {
    LBAudioDetectiveRef d = make_detective(cfg);
    struct LBADShimURL url = { pcm, n };
    LBAudioDetectiveFingerprintRef fp = NULL;
    OSStatus e = LBAudioDetectiveProcessAudioURL(d, (NSURL*)&url, &fp);
    *out_count = dump_fingerprint(fp, out_bits, max_subfps, out_len);
    LBAudioDetectiveFingerprintDispose(fp);
    LBAudioDetectiveDispose(d);
    return e;
}

/* CompareAudioURLs exactly as written (LBAudioDetective.m:442-464). */
int lbad_ref_compare_pcm(const lbad_ref_cfg* cfg, const float* pcm1, SInt64 n1, const float* pcm2, SInt64 n2, UInt32 range, float* out) {
    LBAudioDetectiveRef d = make_detective(cfg);
    struct LBADShimURL u1 = { pcm1, n1 }, u2 = { pcm2, n2 };
    OSStatus e = LBAudioDetectiveCompareAudioURLs(d, (NSURL*)&u1, (NSURL*)&u2, range, out);
    LBAudioDetectiveDispose(d);
    return e;
}

/* Band energies of `n_windows` windows starting at sample `stride*w`, through the reference's own
 * ComputeFrequencies (LBAudioDetective.m:335-408).  The scratch row is 2*bands+64 floats so the LP64
 * memset overrun at m:374-375 (Q14) lands in our buffer and not in the caller's samples. */
int lbad_ref_band_energies(const lbad_ref_cfg* cfg, const float* pcm, SInt64 n, UInt32 n_windows, float* out) {
    LBAudioDetectiveRef d = make_detective(cfg);
    UInt32 N = LBAudioDetectiveGetWindowSize(d), hop = LBAudioDetectiveGetAnalysisStride(d), B = LBAudioDetectiveGetNumberOfPitchSteps(d);
    AudioStreamBasicDescription fmt = LBAudioDetectiveDefaultProcessingFormat();
    fmt.mSampleRate = LBAudioDetectiveGetProcessingSampleRate(d);
    float* win = malloc(sizeof(float) * N);
    float* row = malloc(sizeof(float) * (2 * B + 64));
    OSStatus e = noErr;
    for (UInt32 w = 0; w < n_windows; w++) {
        if ((SInt64)hop * w + N > n) { e = 1; break; }
        memcpy(win, pcm + (size_t)hop * w, sizeof(float) * N);
        e = LBAudioDetectiveComputeFrequencies(d, win, N, fmt, B, row);
        memcpy(out + (size_t)w * B, row, sizeof(float) * B);
    }
    free(win); free(row);
    LBAudioDetectiveDispose(d);
    return e;
}

/* The framing loop of m:250-293 re-driven through the reference's exported internals (ComputeFrequencies,
 * FrameSetRow, SynthesizeFingerprint) with an over-sized band row — the Q14-safe route for window sizes
 * other than 2048/1024.  Optionally dumps the spectral images (before Haar) and Haar coefficients. */
int lbad_ref_process_pcm_direct(const lbad_ref_cfg* cfg, const float* pcm, SInt64 n, Boolean* out_bits, UInt32 max_subfps,
                                UInt32* out_count, UInt32* out_len, float* out_images, float* out_haar) {
    LBAudioDetectiveRef d = make_detective(cfg);
    UInt32 N = LBAudioDetectiveGetWindowSize(d), hop = LBAudioDetectiveGetAnalysisStride(d), B = LBAudioDetectiveGetNumberOfPitchSteps(d);
    UInt32 R = kLBAudioDetectiveDefaultNumberOfRowsPerFrame;
    AudioStreamBasicDescription fmt = LBAudioDetectiveDefaultProcessingFormat();
    fmt.mSampleRate = LBAudioDetectiveGetProcessingSampleRate(d);
    *out_count = 0;
    if (n < (SInt64)N) { LBAudioDetectiveDispose(d); return 1; }
    UInt64 imageWidth = (UInt64)(n - N) / hop, framesCount = imageWidth / R;
    LBAudioDetectiveFrameRef* frames = malloc((framesCount ? framesCount : 1) * sizeof(*frames));
    float* win = malloc(sizeof(float) * N);
    float* row = malloc(sizeof(float) * (2 * B + 64));
    for (UInt64 f = 0; f < framesCount; f++) {
        frames[f] = LBAudioDetectiveFrameNew(R);
        for (UInt32 r = 0; r < R; r++) {
            memcpy(win, pcm + (size_t)hop * (f * R + r), sizeof(float) * N);
            LBAudioDetectiveComputeFrequencies(d, win, N, fmt, B, row);
            LBAudioDetectiveFrameSetRow(frames[f], row, r, B);
            if (out_images && f < max_subfps) memcpy(out_images + ((size_t)f * R + r) * B, row, sizeof(float) * B);
        }
    }
    LBAudioDetectiveFingerprintRef fp = LBAudioDetectiveFingerprintNew(0);
    LBAudioDetectiveSynthesizeFingerprint(d, frames, framesCount, &fp);   /* decomposes the frames in place */
    if (out_haar) for (UInt64 f = 0; f < framesCount && f < max_subfps; f++)
        for (UInt32 r = 0; r < R; r++) memcpy(out_haar + ((size_t)f * R + r) * B, LBAudioDetectiveFrameGetRow(frames[f], r), sizeof(float) * B);
    *out_count = dump_fingerprint(fp, out_bits, max_subfps, out_len);
    LBAudioDetectiveFingerprintDispose(fp);
    for (UInt64 f = 0; f < framesCount; f++) LBAudioDetectiveFrameDispose(frames[f]);
    free(frames); free(win); free(row);
    LBAudioDetectiveDispose(d);
    return 0;
}

/* FrameDecompose (LBAudioDetectiveFrame.m:113-153) on a rows x cols image, in place. */
void lbad_ref_haar(float* image, UInt32 rows, UInt32 cols) {
    LBAudioDetectiveFrameRef f = LBAudioDetectiveFrameNew(rows);
    for (UInt32 r = 0; r < rows; r++) LBAudioDetectiveFrameSetRow(f, image + (size_t)r * cols, r, cols);
    LBAudioDetectiveFrameDecompose(f);
    for (UInt32 r = 0; r < rows; r++) memcpy(image + (size_t)r * cols, LBAudioDetectiveFrameGetRow(f, r), sizeof(float) * cols);
    LBAudioDetectiveFrameDispose(f);
}

/* FrameExtractFingerprint on an ALREADY decomposed rows x cols image; out has 2*t Booleans (zeroed here). */
void lbad_ref_extract_bits(const float* coeffs, UInt32 rows, UInt32 cols, UInt32 t, Boolean* out) {
    LBAudioDetectiveFrameRef f = LBAudioDetectiveFrameNew(rows);
    for (UInt32 r = 0; r < rows; r++) LBAudioDetectiveFrameSetRow(f, (Float32*)coeffs + (size_t)r * cols, r, cols);
    memset(out, 0, 2 * (size_t)t);
    LBAudioDetectiveFrameExtractFingerprint(f, t, out);
    LBAudioDetectiveFrameDispose(f);
}

static LBAudioDetectiveFingerprintRef build_fp(const Boolean* bits, UInt32 count, UInt32 L) {
    LBAudioDetectiveFingerprintRef fp = LBAudioDetectiveFingerprintNew(L);
    for (UInt32 i = 0; i < count; i++) LBAudioDetectiveFingerprintAddSubfingerprint(fp, (Boolean*)bits + (size_t)i * L);
    return fp;
}

float lbad_ref_compare_fp(const Boolean* b1, UInt32 c1, const Boolean* b2, UInt32 c2, UInt32 L, UInt32 range) {
    LBAudioDetectiveFingerprintRef f1 = build_fp(b1, c1, L), f2 = build_fp(b2, c2, L);
    float r = LBAudioDetectiveFingerprintCompareToFingerprint(f1, f2, range);
    LBAudioDetectiveFingerprintDispose(f1); LBAudioDetectiveFingerprintDispose(f2);
    return r;
}

float lbad_ref_compare_sub(const Boolean* s1, const Boolean* s2, UInt32 L, UInt32 range) {
    LBAudioDetectiveFingerprintRef f = LBAudioDetectiveFingerprintNew(L);
    float r = LBAudioDetectiveFingerprintCompareSubfingerprints(f, (Boolean*)s1, (Boolean*)s2, range);
    LBAudioDetectiveFingerprintDispose(f);
    return r;
}

int lbad_ref_set_window_size_status(UInt32 n) {
    LBAudioDetectiveRef d = LBAudioDetectiveNew();
    int e = LBAudioDetectiveSetWindowSize(d, n);
    LBAudioDetectiveDispose(d);
    return e;
}

/* ------------------------------------------------------------------ timing ---- */

static double now_s(void);
void lbad_shim_set_fft_mode(int mode);
int lbad_shim_get_fft_mode(void);

/* seconds per window of the three vDSP calls the reference makes (LBAudioDetective.m:353-355), in the given shim mode */
double lbad_ref_time_fft(UInt32 n, int mode, UInt32 reps) {
    UInt32 log2n = 0; while ((1u << log2n) < n) log2n++;
    FFTSetup setup = vDSP_create_fftsetup(log2n, FFT_RADIX2);
    float* x = malloc(sizeof(float) * n); float* re = malloc(sizeof(float) * n / 2); float* im = malloc(sizeof(float) * n / 2);
    for (UInt32 i = 0; i < n; i++) x[i] = (float)((i * 2654435761u) >> 8) / 16777216.0f - 0.5f;
    COMPLEX_SPLIT A = { re, im };
    int old = lbad_shim_get_fft_mode(); lbad_shim_set_fft_mode(mode);
    volatile float sink = 0;
    double t0 = now_s();
    for (UInt32 r = 0; r < reps; r++) {
        x[r % n] += 1e-3f;
        vDSP_ctoz((COMPLEX*)x, 2, &A, 1, n / 2);
        vDSP_fft_zrip(setup, &A, 1, log2n, FFT_FORWARD);
        vDSP_ztoc(&A, 1, (COMPLEX*)x, 2, n / 2);
        sink += x[7];
        for (UInt32 i = 0; i < 8; i++) x[(r * 8 + i) % n] *= 1e-3f;      /* keep the values bounded */
    }
    double dt = (now_s() - t0) / reps;
    lbad_shim_set_fft_mode(old);
    vDSP_destroy_fftsetup(setup); free(x); free(re); free(im);
    return dt;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

typedef struct {
    const lbad_ref_cfg* cfg; const float* pcm; SInt64 clip_len; UInt32 first, last;
    Boolean* out_bits; UInt32 max_subfps_per_clip; UInt32* out_counts; UInt32 L;
} extract_job;

static void* extract_worker(void* p) {
    extract_job* j = p;
    LBAudioDetectiveRef d = make_detective(j->cfg);     /* one detective per thread (not re-entrant) */
    for (UInt32 c = j->first; c < j->last; c++) {
        struct LBADShimURL url = { j->pcm + (size_t)c * j->clip_len, j->clip_len };
        LBAudioDetectiveFingerprintRef fp = NULL;
        LBAudioDetectiveProcessAudioURL(d, (NSURL*)&url, &fp);
        UInt32 L = 0;
        UInt32 n = j->out_bits ? dump_fingerprint(fp, j->out_bits + (size_t)c * j->max_subfps_per_clip * j->L, j->max_subfps_per_clip, &L)
                               : LBAudioDetectiveFingerprintGetNumberOfSubfingerprints(fp);
        if (j->out_counts) j->out_counts[c] = n;
        LBAudioDetectiveFingerprintDispose(fp);
    }
    LBAudioDetectiveDispose(d);
    return NULL;
}

/* Fingerprint n_clips equal-length clips with `threads` host threads; returns wall seconds.
 * out_bits (optional): [n_clips][max_subfps_per_clip][sublen] Booleans; out_counts (optional): [n_clips]. */
double lbad_ref_extract_batch(const lbad_ref_cfg* cfg, const float* pcm, UInt32 n_clips, SInt64 clip_len, UInt32 threads,
                              Boolean* out_bits, UInt32 max_subfps_per_clip, UInt32* out_counts) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; extract_job jobs[256];
    UInt32 L = cfg && cfg->sublen ? cfg->sublen : 200;
    double t0 = now_s();
    for (UInt32 t = 0; t < threads; t++) {
        jobs[t] = (extract_job){ cfg, pcm, clip_len, (UInt32)((UInt64)n_clips * t / threads), (UInt32)((UInt64)n_clips * (t + 1) / threads),
                                 out_bits, max_subfps_per_clip, out_counts, L };
        pthread_create(&th[t], NULL, extract_worker, &jobs[t]);
    }
    for (UInt32 t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return now_s() - t0;
}

/* Stage dumps of a whole batch on `threads` host threads, every clip through lbad_ref_process_pcm_direct (the reference's own
 * ComputeFrequencies / FrameSetRow / SynthesizeFingerprint; Q14-safe at every window size): out_images and out_haar are
 * [n_clips][subfps_per_clip][128][bands] (either may be NULL), out_bits [n_clips][subfps_per_clip][sublen].  Returns wall seconds. */
typedef struct { const lbad_ref_cfg* cfg; const float* pcm; SInt64 clip_len; UInt32 first, last, per_clip, bands, L; float* img; float* haar; Boolean* bits; } stage_job;

static void* stage_worker(void* p) {
    stage_job* j = p;
    const size_t fsz = (size_t)128 * j->bands;
    for (UInt32 c = j->first; c < j->last; c++) {
        UInt32 cnt = 0, L = 0;
        lbad_ref_process_pcm_direct(j->cfg, j->pcm + (size_t)c * j->clip_len, j->clip_len, j->bits + (size_t)c * j->per_clip * j->L, j->per_clip, &cnt, &L,
                                    j->img ? j->img + (size_t)c * j->per_clip * fsz : NULL, j->haar ? j->haar + (size_t)c * j->per_clip * fsz : NULL);
    }
    return NULL;
}

double lbad_ref_extract_batch_stages(const lbad_ref_cfg* cfg, const float* pcm, UInt32 n_clips, SInt64 clip_len, UInt32 threads,
                                     UInt32 subfps_per_clip, float* out_images, float* out_haar, Boolean* out_bits) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > n_clips) threads = n_clips ? n_clips : 1;
    pthread_t th[256]; stage_job jobs[256];
    double t0 = now_s();
    for (UInt32 t = 0; t < threads; t++) {
        jobs[t] = (stage_job){ cfg, pcm, clip_len, (UInt32)((UInt64)n_clips * t / threads), (UInt32)((UInt64)n_clips * (t + 1) / threads), subfps_per_clip,
                               cfg->bands ? cfg->bands : 32, cfg->sublen ? cfg->sublen : 200, out_images, out_haar, out_bits };
        pthread_create(&th[t], NULL, stage_worker, &jobs[t]);
    }
    for (UInt32 t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return now_s() - t0;
}

typedef struct {
    LBAudioDetectiveFingerprintRef* db; LBAudioDetectiveFingerprintRef* q; UInt32 n_q, first, last, range; float* scores; UInt32 n_db;
} search_job;

static void* search_worker(void* p) {
    search_job* j = p;
    for (UInt32 c = j->first; c < j->last; c++)
        for (UInt32 q = 0; q < j->n_q; q++)
            j->scores[(size_t)q * j->n_db + c] = LBAudioDetectiveFingerprintCompareToFingerprint(j->db[c], j->q[q], j->range);
    return NULL;
}

/* Linear scan: scores[q][c] = CompareToFingerprint(db[c], query[q], range) (archive first, query second, as
 * LBAudioDetectiveTests.m:68 does).  db_bits: [n_db][db_count][L], q_bits: [n_q][q_count][L].  Returns wall
 * seconds of the compare loop only (fingerprint construction excluded). */
double lbad_ref_search(const Boolean* db_bits, UInt32 n_db, UInt32 db_count, const Boolean* q_bits, UInt32 n_q, UInt32 q_count,
                       UInt32 L, UInt32 range, UInt32 threads, float* scores) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    LBAudioDetectiveFingerprintRef* db = malloc(sizeof(*db) * (n_db ? n_db : 1));
    LBAudioDetectiveFingerprintRef* q = malloc(sizeof(*q) * (n_q ? n_q : 1));
    for (UInt32 i = 0; i < n_db; i++) db[i] = build_fp(db_bits + (size_t)i * db_count * L, db_count, L);
    for (UInt32 i = 0; i < n_q; i++) q[i] = build_fp(q_bits + (size_t)i * q_count * L, q_count, L);
    pthread_t th[256]; search_job jobs[256];
    double t0 = now_s();
    for (UInt32 t = 0; t < threads; t++) {
        jobs[t] = (search_job){ db, q, n_q, (UInt32)((UInt64)n_db * t / threads), (UInt32)((UInt64)n_db * (t + 1) / threads), range, scores, n_db };
        pthread_create(&th[t], NULL, search_worker, &jobs[t]);
    }
    for (UInt32 t = 0; t < threads; t++) pthread_join(th[t], NULL);
    double dt = now_s() - t0;
    for (UInt32 i = 0; i < n_db; i++) LBAudioDetectiveFingerprintDispose(db[i]);
    for (UInt32 i = 0; i < n_q; i++) LBAudioDetectiveFingerprintDispose(q[i]);
    free(db); free(q);
    return dt;
}
