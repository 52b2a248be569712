/* TEST INFRASTRUCTURE — not part of the product.
 *
 * CPU restatement ("port") of the reference's fingerprint path in plain C, used only as the checker for the
 * CUDA path and as a timed CPU baseline.  Parity status: PINNED against the compiled reference
 * (oracle/_ref/libLBAudioDetectiveRef.so, built from /root/reference by oracle/Makefile) — tests/test_oracle_vs_ref.py
 * runs both on identical PCM and demands bit-identical band energies, Haar coefficients, Booleans and scores — and
 * against tests/golden/ (outputs of that compiled reference, committed with the generating script).
 * The two Apple-closed boundaries stay pinned BY DEFINITION only: vDSP_fft_zrip := exact real DFT
 * x2 rounded to f32; NSMutableArray sort := stable.
 *
 * Citations: m: = LBAudioDetective/LBAudioDetective.m, Frame.m / FP.m likewise, all under /root/reference.
 * Compile with -ffp-contract=off: the reference's float expressions are evaluated op by op in f32.
 */
#include "lbad_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

void lbad_oracle_default_cfg(lbad_oracle_cfg* c) {           /* m:22-26, m:128 */
    c->window = 2048; c->stride = 64; c->bands = 32; c->sublen = 200; c->sample_rate = 5512.0;
}

/* ---------------------------------------------------------------- band table ---- */

void lbad_oracle_band_table(const lbad_oracle_cfg* c, uint32_t nframes, uint32_t* idx, uint32_t* klow, uint32_t* khigh) {
    uint32_t B = c->bands;
    double maxFreq = c->sample_rate / 2.0;                                   /* m:362 */
    double minFreq = 318.0;                                                  /* m:363 */
    double logBase = exp(log(maxFreq / minFreq) / B);                        /* m:365 */
    double mincoef = (double)c->window / c->sample_rate * minFreq;           /* m:366 */
    for (uint32_t j = 0; j <= B; j++) {                                      /* m:368-371 */
        uint32_t start = (uint32_t)((pow(logBase, j) - 1.0) * mincoef);
        idx[j] = start + (uint32_t)mincoef;
    }
    for (uint32_t i = 0; i < B; i++) {                                       /* m:382-383 */
        klow[i]  = (uint32_t)(((2 * idx[i])     / (c->sample_rate / nframes)) - 1);
        khigh[i] = (uint32_t)(((2 * idx[i + 1]) / (c->sample_rate / nframes)) - 1);
    }
}

/* ----------------------------------------------------------------------- FFT ---- */

typedef struct { uint32_t n; double* tw; float* twf; double* a; double* b; float* fa; float* fb; } fft_plan;

static fft_plan* plan_new(uint32_t n) {
    fft_plan* p = calloc(1, sizeof *p);
    uint32_t h = n / 2 ? n / 2 : 1;
    p->n = n;
    p->tw = malloc(sizeof(double) * 2 * h); p->twf = malloc(sizeof(float) * 2 * h);
    for (uint32_t j = 0; j < h; j++) {
        double th = 2.0 * M_PI * (double)j / (double)n;
        p->tw[2*j] = cos(th); p->tw[2*j+1] = sin(th);
        p->twf[2*j] = (float)p->tw[2*j]; p->twf[2*j+1] = (float)p->tw[2*j+1];
    }
    p->a = malloc(sizeof(double) * 2 * h); p->b = malloc(sizeof(double) * 2 * h);
    p->fa = malloc(sizeof(float) * 2 * h); p->fb = malloc(sizeof(float) * 2 * h);
    return p;
}
static void plan_free(fft_plan* p) { if (!p) return; free(p->tw); free(p->twf); free(p->a); free(p->b); free(p->fa); free(p->fb); free(p); }

/* M-point complex forward DFT (e^{-i theta}), Stockham radix-2 DIF, double.  Same operation order as the
 * shim's f64 kernel so that oracle and compiled reference agree to the last bit. */
static double* cfft_f64(double* x, double* y, uint32_t M, const double* tw) {
    uint32_t N = 2 * M;
    for (uint32_t n = M, s = 1; n > 1; n >>= 1, s <<= 1) {
        uint32_t m = n / 2, step = N / n;
        for (uint32_t p = 0; p < m; p++) {
            double wr = tw[2 * p * step], wi = -tw[2 * p * step + 1];
            const double* xa = x + 2 * (size_t)s * p; const double* xb = x + 2 * (size_t)s * (p + m);
            double* ya = y + 2 * (size_t)s * (2 * p); double* yb = y + 2 * (size_t)s * (2 * p + 1);
            for (uint32_t q = 0; q < s; q++) {
                double ar = xa[2*q], ai = xa[2*q+1], br = xb[2*q], bi = xb[2*q+1];
                double dr = ar - br, di = ai - bi;
                ya[2*q] = ar + br;           ya[2*q+1] = ai + bi;
                yb[2*q] = dr * wr - di * wi; yb[2*q+1] = dr * wi + di * wr;
            }
        }
        double* t = x; x = y; y = t;
    }
    return x;
}
static float* cfft_f32(float* x, float* y, uint32_t M, const float* tw) {
    uint32_t N = 2 * M;
    for (uint32_t n = M, s = 1; n > 1; n >>= 1, s <<= 1) {
        uint32_t m = n / 2, step = N / n;
        for (uint32_t p = 0; p < m; p++) {
            float wr = tw[2 * p * step], wi = -tw[2 * p * step + 1];
            const float* xa = x + 2 * (size_t)s * p; const float* xb = x + 2 * (size_t)s * (p + m);
            float* ya = y + 2 * (size_t)s * (2 * p); float* yb = y + 2 * (size_t)s * (2 * p + 1);
            for (uint32_t q = 0; q < s; q++) {
                float ar = xa[2*q], ai = xa[2*q+1], br = xb[2*q], bi = xb[2*q+1];
                float dr = ar - br, di = ai - bi;
                ya[2*q] = ar + br;           ya[2*q+1] = ai + bi;
                yb[2*q] = dr * wr - di * wi; yb[2*q+1] = dr * wi + di * wr;
            }
        }
        float* t = x; x = y; y = t;
    }
    return x;
}

/* m:353-355 under the vDSP definition (Q2): pack pairs as complex (ctoz), real-FFT with x2 scaling, unpack (ztoc).
 * out[2k] = 2 Re X[k], out[2k+1] = 2 Im X[k] for 0 < k < N/2; out[0] = 2 X[0]; out[1] = 2 X[N/2]. */
static void fft2x_with_plan(fft_plan* p, const float* x, float* out, int f32) {
    uint32_t N = p->n, M = N / 2;
    if (!f32) {
        for (uint32_t i = 0; i < N; i++) p->a[i] = x[i];
        const double* Z = cfft_f64(p->a, p->b, M, p->tw);
        double dc = 2.0 * (Z[0] + Z[1]), ny = 2.0 * (Z[0] - Z[1]);
        for (uint32_t k = 1; k < M; k++) {
            double zr = Z[2*k], zi = Z[2*k+1], yr = Z[2*(M-k)], yi = -Z[2*(M-k)+1];
            double er = zr + yr, ei = zi + yi, dr = zr - yr, di = zi - yi;
            double c = p->tw[2*k], s = p->tw[2*k+1];
            double tr = -s * dr + c * di, ti = -c * dr - s * di;
            out[2*k] = (float)(er + tr); out[2*k+1] = (float)(ei + ti);
        }
        out[0] = (float)dc; out[1] = (float)ny;
    } else {
        for (uint32_t i = 0; i < N; i++) p->fa[i] = x[i];
        const float* Z = cfft_f32(p->fa, p->fb, M, p->twf);
        float dc = 2.0f * (Z[0] + Z[1]), ny = 2.0f * (Z[0] - Z[1]);
        for (uint32_t k = 1; k < M; k++) {
            float zr = Z[2*k], zi = Z[2*k+1], yr = Z[2*(M-k)], yi = -Z[2*(M-k)+1];
            float er = zr + yr, ei = zi + yi, dr = zr - yr, di = zi - yi;
            float c = p->twf[2*k], s = p->twf[2*k+1];
            float tr = -s * dr + c * di, ti = -c * dr - s * di;
            out[2*k] = er + tr; out[2*k+1] = ei + ti;
        }
        out[0] = dc; out[1] = ny;
    }
}

void lbad_oracle_fft2x(const float* x, uint32_t n, float* out) {
    fft_plan* p = plan_new(n); fft2x_with_plan(p, x, out, 0); plan_free(p);
}

/* -------------------------------------------------------------- band energy ---- */

typedef struct { lbad_oracle_cfg cfg; fft_plan* plan; uint32_t* idx; uint32_t* klow; uint32_t* khigh; float* spec; int f32; } ctx_t;

static ctx_t* ctx_new(const lbad_oracle_cfg* c, int f32) {
    ctx_t* x = calloc(1, sizeof *x);
    x->cfg = *c; x->f32 = f32;
    x->plan = plan_new(c->window);
    x->idx = malloc(sizeof(uint32_t) * (c->bands + 1)); x->klow = malloc(sizeof(uint32_t) * c->bands); x->khigh = malloc(sizeof(uint32_t) * c->bands);
    lbad_oracle_band_table(c, c->window, x->idx, x->klow, x->khigh);   /* the reference recomputes this per window (m:361-371) */
    x->spec = malloc(sizeof(float) * (c->window + 2));
    return x;
}
static void ctx_free(ctx_t* x) { plan_free(x->plan); free(x->idx); free(x->klow); free(x->khigh); free(x->spec); free(x); }

static void window_bands(ctx_t* x, const float* win, float* out) {
    const lbad_oracle_cfg* c = &x->cfg;
    fft2x_with_plan(x->plan, win, x->spec, x->f32);                          /* m:351-355 */
    const float* samples = x->spec;
    uint32_t width = (uint32_t)(c->window / 2.0);                            /* m:373 */
    float scale = (float)(width / 2);                                        /* m:391, integer division */
    for (uint32_t i = 0; i < c->bands; i++) {                                /* m:379 */
        float p = 0.0f;                                                      /* m:384 */
        for (uint32_t k = x->klow[i]; k < x->khigh[i]; k++) {                /* m:386 */
            float re = samples[2 * k], img = samples[2 * k + 1];             /* m:387-388 */
            if (re > 0.0) re /= scale;                                       /* m:390-392 (Q4: positive parts only) */
            if (img > 0.0) img /= scale;                                     /* m:393-395 */
            float v = (re * re) + (img * img);                               /* m:397 */
            if (v == v && isfinite(v)) p += v;                               /* m:398-401 */
        }
        out[i] = p / (float)(x->idx[i + 1] - x->idx[i]);                     /* m:404 (Q5: first-table divisor) */
    }
}

void lbad_oracle_window_bands(const lbad_oracle_cfg* c, const float* win, float* out) {
    ctx_t* x = ctx_new(c, 0); window_bands(x, win, out); ctx_free(x);
}

/* ---------------------------------------------------------------------- Haar ---- */

static void haar_1d(float* a, uint32_t n, float* tmp) {                      /* Frame.m:134-153 */
    for (uint32_t i = 0; i < n; i++) a[i] /= sqrtf(n);                       /* Frame.m:137-139 */
    while (n > 1) {                                                          /* Frame.m:143 */
        n /= 2;
        for (uint32_t i = 0; i < n; i++) {                                   /* Frame.m:145-148 */
            tmp[i]     = ((a[2 * i] + a[2 * i + 1]) / sqrtf(2.0f));
            tmp[n + i] = ((a[2 * i] - a[2 * i + 1]) / sqrtf(2.0f));
        }
        for (uint32_t i = 0; i < 2 * n; i++) a[i] = tmp[i];                  /* Frame.m:149-151 */
    }
}

void lbad_oracle_haar(float* image, uint32_t rows, uint32_t cols) {          /* Frame.m:113-132 */
    uint32_t m = rows > cols ? rows : cols;
    float* tmp = malloc(sizeof(float) * (m ? m : 1)); float* col = malloc(sizeof(float) * (rows ? rows : 1));
    for (uint32_t r = 0; r < rows; r++) haar_1d(image + (size_t)r * cols, cols, tmp);   /* Frame.m:114-116 */
    for (uint32_t c = 0; c < cols; c++) {                                               /* Frame.m:118-131 */
        for (uint32_t r = 0; r < rows; r++) col[r] = image[(size_t)r * cols + c];
        haar_1d(col, rows, tmp);
        for (uint32_t r = 0; r < rows; r++) image[(size_t)r * cols + c] = col[r];
    }
    free(tmp); free(col);
}

/* --------------------------------------------------------------------- top-t ---- */

void lbad_oracle_extract_bits(const float* v, uint32_t n, uint32_t t, uint8_t* out) {   /* Frame.m:165-191 */
    uint32_t* idx = malloc(sizeof(uint32_t) * (n ? n : 1)); uint32_t* tmp = malloc(sizeof(uint32_t) * (n ? n : 1));
    for (uint32_t i = 0; i < n; i++) idx[i] = i;                             /* Frame.m:170-174, flat = row*rowLength+col */
    for (uint32_t w = 1; w < n; w *= 2) {                                    /* Frame.m:176-178 as a STABLE descending |v| sort */
        for (uint32_t lo = 0; lo < n; lo += 2 * w) {
            uint32_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, a = lo, b = mid, o = lo;
            while (a < mid && b < hi) tmp[o++] = (fabs((double)v[idx[b]]) > fabs((double)v[idx[a]])) ? idx[b++] : idx[a++];
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, sizeof(uint32_t) * n);
    }
    memset(out, 0, 2 * (size_t)t);
    for (uint32_t i = 0; i < t && i < n; i++) {                              /* Frame.m:182-190 */
        double value = v[idx[i]];
        if (value > 0.0) out[2 * i] = 1; else if (value < 0.0) out[2 * i + 1] = 1;
    }
    free(idx); free(tmp);
}

/* ------------------------------------------------------------------- process ---- */

uint64_t lbad_oracle_subfp_count(const lbad_oracle_cfg* c, int64_t n) {      /* m:250-255 */
    if (n < (int64_t)c->window || c->stride == 0) return 0;                  /* the reference underflows here (Q7); rejected */
    uint64_t imageWidth = (uint64_t)(n - c->window) / c->stride;
    return imageWidth / LBAD_ORACLE_ROWS_PER_FRAME;
}

static int process_ctx(ctx_t* x, const float* pcm, int64_t n, uint8_t* out_bits, uint32_t max_subfps, uint32_t* out_count,
                       float* out_images, float* out_haar) {
    const lbad_oracle_cfg* c = &x->cfg;
    const uint32_t R = LBAD_ORACLE_ROWS_PER_FRAME, B = c->bands, L = c->sublen;
    uint64_t frames = lbad_oracle_subfp_count(c, n);
    *out_count = (uint32_t)frames;
    float* image = malloc(sizeof(float) * R * B);
    uint8_t* sub = malloc(2 * (size_t)L + 2);
    for (uint64_t f = 0; f < frames && f < max_subfps; f++) {
        for (uint32_t r = 0; r < R; r++)                                     /* m:262-290: window i starts at stride*i */
            window_bands(x, pcm + (size_t)c->stride * (f * R + r), image + (size_t)r * B);
        if (out_images) memcpy(out_images + (size_t)f * R * B, image, sizeof(float) * R * B);
        lbad_oracle_haar(image, R, B);                                       /* m:320 */
        if (out_haar) memcpy(out_haar + (size_t)f * R * B, image, sizeof(float) * R * B);
        lbad_oracle_extract_bits(image, R * B, L, sub);                      /* m:321-324: 2L Booleans written ... */
        if (out_bits) memcpy(out_bits + (size_t)f * L, sub, L);              /* m:326-328, FP.m:92-94: ... L kept (Q10) */
    }
    free(image); free(sub);
    return 0;
}

int lbad_oracle_process(const lbad_oracle_cfg* c, const float* pcm, int64_t n, uint8_t* out_bits, uint32_t max_subfps,
                        uint32_t* out_count, float* out_images, float* out_haar) {
    ctx_t* x = ctx_new(c, 0);
    int e = process_ctx(x, pcm, n, out_bits, max_subfps, out_count, out_images, out_haar);
    ctx_free(x);
    return e;
}

/* ------------------------------------------------------------------- compare ---- */

float lbad_oracle_compare_sub(const uint8_t* s1, const uint8_t* s2, uint32_t len1, uint32_t range) {   /* FP.m:151-176 */
    uint32_t possible = 0, hits = 0;
    uint32_t lim = range < len1 ? range : len1;                              /* FP.m:155 */
    for (uint32_t i = 0; i < lim; i += 2) {
        uint8_t a1 = s1[i], a2 = s1[i + 1];
        if (a1 || a2) {                                                      /* FP.m:159 */
            possible++;
            uint8_t b1 = s2[i], b2 = s2[i + 1];
            if ((a1 == b1) && (a2 == b2)) hits++;                            /* FP.m:165 */
        }
    }
    if (possible <= 0) return 0.0f;                                          /* FP.m:171-173 */
    return (float)hits / (float)possible;                                    /* FP.m:175 */
}

float lbad_oracle_compare_fp(const uint8_t* b1, uint32_t c1, uint32_t l1, const uint8_t* b2, uint32_t c2, uint32_t l2, uint32_t range) {
    if (c1 < c2) {                                                           /* FP.m:123-131: swap so fp1 has MORE subfps */
        const uint8_t* tb = b1; b1 = b2; b2 = tb;
        uint32_t t = c1; c1 = c2; c2 = t;
        t = l1; l1 = l2; l2 = t;
    }
    float match = 0.0f;                                                      /* FP.m:133 */
    for (uint32_t offset = 0; offset <= c1 - c2; offset++) {                 /* FP.m:136 */
        float sum = 0.0f;
        for (uint32_t i = 0; i < c2; i++)                                    /* FP.m:139-142 */
            sum += lbad_oracle_compare_sub(b1 + (size_t)(i + offset) * l1, b2 + (size_t)i * l2, l1, range);
        float mean = sum / (float)c2;                                        /* FP.m:144 */
        match = (match < mean) ? mean : match;                               /* Apple MAX: keeps `match` when mean is NaN (c2 == 0) */
    }
    return match;
}

int lbad_oracle_compare_pcm(const lbad_oracle_cfg* c, const float* p1, int64_t n1, const float* p2, int64_t n2, uint32_t range, float* out) {
    if (range == 0) range = c->sublen;                                       /* m:443-445 */
    uint32_t c1 = (uint32_t)lbad_oracle_subfp_count(c, n1), c2 = (uint32_t)lbad_oracle_subfp_count(c, n2), k;
    uint8_t* b1 = calloc((size_t)(c1 ? c1 : 1) * c->sublen, 1); uint8_t* b2 = calloc((size_t)(c2 ? c2 : 1) * c->sublen, 1);
    lbad_oracle_process(c, p1, n1, b1, c1, &k, NULL, NULL);                  /* m:449 */
    lbad_oracle_process(c, p2, n2, b2, c2, &k, NULL, NULL);                  /* m:453 */
    *out = lbad_oracle_compare_fp(b1, c1, c->sublen, b2, c2, c->sublen, range);   /* m:457 */
    free(b1); free(b2);
    return 0;
}

/* --------------------------------------------------------- batch / timing ---- */

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

typedef struct { const lbad_oracle_cfg* c; const float* pcm; int64_t clip_len; uint32_t first, last; uint8_t* out; uint32_t maxs; uint32_t* counts; int f32; } ejob;
static void* eworker(void* p) {
    ejob* j = p; ctx_t* x = ctx_new(j->c, j->f32);
    for (uint32_t i = j->first; i < j->last; i++) {
        uint32_t cnt = 0;
        process_ctx(x, j->pcm + (size_t)i * j->clip_len, j->clip_len, j->out ? j->out + (size_t)i * j->maxs * j->c->sublen : NULL,
                    j->out ? j->maxs : 0xffffffffu, &cnt, NULL, NULL);
        if (j->counts) j->counts[i] = cnt;
    }
    ctx_free(x); return NULL;
}
double lbad_oracle_extract_batch(const lbad_oracle_cfg* c, const float* pcm, uint32_t n_clips, int64_t clip_len, uint32_t threads,
                                 uint8_t* out_bits, uint32_t max_subfps_per_clip, uint32_t* out_counts, int fft_f32) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; ejob jobs[256];
    double t0 = now_s();
    for (uint32_t t = 0; t < threads; t++) {
        jobs[t] = (ejob){ c, pcm, clip_len, (uint32_t)((uint64_t)n_clips * t / threads), (uint32_t)((uint64_t)n_clips * (t + 1) / threads),
                          out_bits, max_subfps_per_clip, out_counts, fft_f32 };
        pthread_create(&th[t], NULL, eworker, &jobs[t]);
    }
    for (uint32_t t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return now_s() - t0;
}

typedef struct { const uint8_t* db; uint32_t n_db, dbc; const uint8_t* q; uint32_t n_q, qc, L, range, first, last; float* scores; } sjob;
static void* sworker(void* p) {
    sjob* j = p;
    for (uint32_t c = j->first; c < j->last; c++)
        for (uint32_t q = 0; q < j->n_q; q++)
            j->scores[(size_t)q * j->n_db + c] = lbad_oracle_compare_fp(j->db + (size_t)c * j->dbc * j->L, j->dbc, j->L,
                                                                        j->q + (size_t)q * j->qc * j->L, j->qc, j->L, j->range);
    return NULL;
}
double lbad_oracle_search(const uint8_t* db_bits, uint32_t n_db, uint32_t db_count, const uint8_t* q_bits, uint32_t n_q, uint32_t q_count,
                          uint32_t L, uint32_t range, uint32_t threads, float* scores) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; sjob jobs[256];
    double t0 = now_s();
    for (uint32_t t = 0; t < threads; t++) {
        jobs[t] = (sjob){ db_bits, n_db, db_count, q_bits, n_q, q_count, L, range,
                          (uint32_t)((uint64_t)n_db * t / threads), (uint32_t)((uint64_t)n_db * (t + 1) / threads), scores };
        pthread_create(&th[t], NULL, sworker, &jobs[t]);
    }
    for (uint32_t t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return now_s() - t0;
}

/* ------------------------------------------------------------- synthetic PCM ---- */

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static double u01(uint64_t h) { return (double)(h >> 40) * (1.0 / 16777216.0); }   /* top 24 bits -> [0,1) */

/* x[n] = 0.5 sin(phi[n]) + 0.2 sin(2 pi f_t n / sr) + 0.1 u[n], clipped to [-1,1].
 * phi = phase of a linear chirp f0 -> f1 over the clip; f0 in [250,600], f1 in [1400,2000], f_t in [400,1800] Hz. */
void lbad_synth_clip(uint64_t base_seed, uint64_t clip_id, int64_t n, double sr, float* out) {
    uint64_t s = splitmix64(base_seed ^ splitmix64(clip_id));
    double f0 = 250.0 + 350.0 * u01(splitmix64(s + 1));
    double f1 = 1400.0 + 600.0 * u01(splitmix64(s + 2));
    double ft = 400.0 + 1400.0 * u01(splitmix64(s + 3));
    double T = (double)n / sr, kr = (f1 - f0) / (T > 0 ? T : 1.0);
    uint64_t ns = splitmix64(s + 4);
    for (int64_t i = 0; i < n; i++) {
        double t = (double)i / sr;
        double phi = 2.0 * M_PI * (f0 * t + 0.5 * kr * t * t);
        double u = u01(splitmix64(ns + (uint64_t)i)) - 0.5;
        double v = 0.5 * sin(phi) + 0.2 * sin(2.0 * M_PI * ft * t) + 0.1 * u;
        if (v > 1.0) v = 1.0;
        if (v < -1.0) v = -1.0;
        out[i] = (float)v;
    }
}

/* uniform noise of the given full-scale amplitude (essay p.34-35 uses 1.58 % and 3.16 %), clipped */
void lbad_synth_add_noise(uint64_t seed, int64_t n, double amplitude, float* io) {
    uint64_t ns = splitmix64(seed ^ 0xA5A5A5A5DEADBEEFull);
    for (int64_t i = 0; i < n; i++) {
        double v = (double)io[i] + amplitude * 2.0 * (u01(splitmix64(ns + (uint64_t)i)) - 0.5);
        if (v > 1.0) v = 1.0;
        if (v < -1.0) v = -1.0;
        io[i] = (float)v;
    }
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Recording-rate -> processing-rate conversion.  The reference delegates this to ExtAudioFile's client format
 * (LBAudioDetective.m:229, m:275 — Apple's converter, closed), so there is nothing of the reference to restate: this is the
 * scalar restatement of the DEFINITION in include/LBAudioDetectiveResample.h, written from that text (PARITY UNPINNED against
 * Apple; the CUDA kernel is checked against this bit for bit, and this against scipy / analytic tones in tests/).
 * ------------------------------------------------------------------------------------------------------------------- */
static double rs_i0(double x) {
    double s = 1.0, t = 1.0;
    for (int k = 1; k < 200; k++) { t *= (0.5 * x / k) * (0.5 * x / k); s += t; if (t < 1e-20 * s) break; }
    return s;
}
static double rs_kaiser(double x, double beta) { return fabs(x) >= 1.0 ? 0.0 : rs_i0(beta * sqrt(1.0 - x * x)) / rs_i0(beta); }
static double rs_sinc(double x) { return x == 0.0 ? 1.0 : sin(M_PI * x) / (M_PI * x); }

uint64_t lbad_oracle_resampled_length(double in_rate, double out_rate, uint64_t n_in) {
    return (uint64_t)floor((double)n_in * out_rate / in_rate + 1e-9);
}

/* out must hold lbad_oracle_resampled_length() samples; returns that count (0 on unsupported rates) */
uint64_t lbad_oracle_resample(double in_rate, double out_rate, const float* x, int64_t n_in, float* out) {
    if (!(out_rate > 0.0) || !(in_rate >= out_rate)) return 0;
    const double rho = in_rate / out_rate;
    int64_t D = (int64_t)floor(rho / 2.0); if (D < 1) D = 1;
    const double rho2 = rho / (double)D;
    const int64_t H1 = 6 * D, T1 = 2 * H1 + 1;
    float* g = NULL;
    if (D > 1) {                                                     /* stage-1 taps */
        g = malloc(T1 * sizeof(float)); double* t = malloc(T1 * sizeof(double)); double sum = 0.0;
        for (int64_t i = 0; i < T1; i++) { const double u = (double)(i - H1); t[i] = rs_sinc(u / (double)D) / (double)D * rs_kaiser(u / (double)(H1 + 1), 8.0); sum += t[i]; }
        for (int64_t i = 0; i < T1; i++) g[i] = (float)(t[i] / sum);
        free(t);
    }
    const double gamma = 0.97 / rho2;
    const int64_t H2 = (int64_t)ceil(10.0 / gamma), T2 = 2 * H2;
    float* hc = malloc((size_t)(65 * T2) * sizeof(float)); double* row = malloc(T2 * sizeof(double));
    for (int p = 0; p <= 64; p++) {                                  /* coarse-phase rows of stage 2 */
        double sum = 0.0;
        for (int64_t i = 0; i < T2; i++) {
            const double u = (double)(i - H2 + 1) - (double)p / 64.0;
            row[i] = fabs(u) < (double)H2 ? gamma * rs_sinc(gamma * u) * rs_kaiser(u / (double)H2, 9.0) : 0.0;
            sum += row[i];
        }
        for (int64_t i = 0; i < T2; i++) hc[(size_t)p * T2 + i] = (float)(row[i] / sum);
    }
    free(row);
    const uint64_t n_out = lbad_oracle_resampled_length(in_rate, out_rate, (uint64_t)n_in);
    for (uint64_t m = 0; m < n_out; m++) {
        const double pos = (double)m * rho2, fl = floor(pos), fp = (pos - fl) * 64.0;
        const int p = (int)fp; const float a = (float)(fp - (double)p);
        const int64_t i0 = (int64_t)fl;
        float s0 = 0.0f, s1 = 0.0f;
        for (int64_t i = 0; i < T2; i++) {
            const int64_t n = i0 + i - H2 + 1;                       /* index into the stage-1 sequence */
            float y;
            if (D > 1) {
                y = 0.0f;                                            /* taps in polyphase order: t = p, p + D, p + 2 D, ... for p = 0 .. D - 1 */
                for (int64_t p = 0; p < D; p++)
                    for (int64_t t = p; t < T1; t += D) { const int64_t k = D * n + t - H1; y = fmaf(g[t], (k >= 0 && k < n_in) ? x[k] : 0.0f, y); }
            } else y = (n >= 0 && n < n_in) ? x[n] : 0.0f;
            s0 = fmaf(hc[(size_t)p * T2 + i], y, s0);
            s1 = fmaf(hc[(size_t)(p + 1) * T2 + i], y, s1);
        }
        out[m] = fmaf(a, s1 - s0, s0);
    }
    free(g); free(hc);
    return n_out;
}
