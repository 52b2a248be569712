/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Link-time stand-ins for the Apple system calls the reference makes on the fingerprint path, so the
 * reference's own .m files (compiled in place from /root/reference by oracle/Makefile) run on Linux:
 *
 *   - Accelerate vDSP: vDSP_create_fftsetup / vDSP_destroy_fftsetup / vDSP_ctoz / vDSP_fft_zrip / vDSP_ztoc
 *     (call sites LBAudioDetective.m:106,179,192,353-355).  Apple's source is closed; the shim DEFINES
 *     vDSP_fft_zrip as its documented contract: forward real DFT with e^{-i theta}, scaled by 2, DC in
 *     realp[0], Nyquist packed into imagp[0].  Two interchangeable kernels:
 *       mode 0 "f64"  (default, used for parity): the DFT evaluated in double precision and rounded once
 *                     to float — the centre of the "within f32 rounding" ball both vDSP and the GPU live in;
 *       mode 1 "f32"  a float Stockham radix-2 FFT with table twiddles (round 1's timing stand-in; kept for
 *                     comparison — it costs about 12 us per 2048-point window);
 *       mode 2 "fast" (used for the CPU timing baseline, bench.py --impl reference and cpu_baseline): a
 *                     single-precision four-step FFT (32 x R points, split re/im arrays, every loop a
 *                     contiguous SIMD loop; AVX-512 / AVX2 clones picked at load time) — what a tuned vDSP
 *                     costs, so that the GPU/CPU ratio is not inflated by a slow stand-in.
 *     Select with lbad_shim_set_fft_mode() or the LBAD_SHIM_FFT=f64|f32|fast environment variable.
 *   - AudioToolbox ExtAudioFile*: a memory-backed "file".  The NSURL* the reference passes around is a
 *     pointer to {const float* samples; int64 count}; Read copies from a cursor, Seek sets the cursor
 *     (LBAudioDetective.m:224-237,275,288).  Decode/resample is out of scope.
 *   - AudioConverter*: aborting stubs (dead code in the reference, LBAudioDetective.m:340-347,413-437).
 */
#include <Foundation/Foundation.h>
#include <AudioToolbox/AudioToolbox.h>
#include <Accelerate/Accelerate.h>

/* ------------------------------------------------------------------ vDSP ---- */

struct LBADShimFFTSetup {
    unsigned log2n;      /* of the REAL transform length N */
    size_t   n;          /* N */
    double*  twd;        /* cos/sin(2*pi*j/N), j < N/2, interleaved — f64 kernel + split pass */
    float*   twf;        /* same in float — f32 kernel */
    double*  wa; double* wb;   /* f64 ping-pong, N/2 complex each */
    float*   fa; float*  fb;   /* f32 ping-pong */
    struct LBADFastFFT* fast;  /* mode 2 tables and scratch */
};
struct LBADFastFFT;
static struct LBADFastFFT* fast_new(size_t n);
static void fast_free(struct LBADFastFFT* f);
static void fast_zrip(struct LBADFastFFT* f, float* re, float* im);

static int g_fft_mode = -1;

void lbad_shim_set_fft_mode(int mode) { g_fft_mode = mode == 2 ? 2 : mode ? 1 : 0; }
int lbad_shim_get_fft_mode(void) {
    if (g_fft_mode < 0) {
        const char* e = getenv("LBAD_SHIM_FFT");
        g_fft_mode = (e && strcmp(e, "f32") == 0) ? 1 : (e && strcmp(e, "fast") == 0) ? 2 : 0;
    }
    return g_fft_mode;
}

FFTSetup vDSP_create_fftsetup(vDSP_Length log2n, FFTRadix radix) {
    (void)radix;
    struct LBADShimFFTSetup* s = calloc(1, sizeof *s);
    s->log2n = (unsigned)log2n;
    s->n = (size_t)1 << log2n;
    size_t h = s->n / 2 ? s->n / 2 : 1;
    s->twd = malloc(2 * h * sizeof(double));
    s->twf = malloc(2 * h * sizeof(float));
    for (size_t j = 0; j < h; j++) {
        double a = 2.0 * M_PI * (double)j / (double)s->n;
        s->twd[2*j] = cos(a); s->twd[2*j+1] = sin(a);
        s->twf[2*j] = (float)s->twd[2*j]; s->twf[2*j+1] = (float)s->twd[2*j+1];
    }
    s->wa = malloc(2 * h * sizeof(double)); s->wb = malloc(2 * h * sizeof(double));
    s->fa = malloc(2 * h * sizeof(float));  s->fb = malloc(2 * h * sizeof(float));
    s->fast = fast_new(s->n);
    return s;
}

void vDSP_destroy_fftsetup(FFTSetup s) {
    if (!s) return;   /* the reference destroys a NULL setup on its first SetWindowSize (LBAudioDetective.m:179) */
    fast_free(s->fast); free(s->twd); free(s->twf); free(s->wa); free(s->wb); free(s->fa); free(s->fb); free(s);
}

void vDSP_ctoz(const DSPComplex* C, vDSP_Stride IC, const DSPSplitComplex* Z, vDSP_Stride IZ, vDSP_Length N) {
    /* IC is in floats (2 = contiguous complex), as in vDSP */
    const float* c = (const float*)C;
    if (IC == 2 && IZ == 1) {            /* the reference's call (m:353): contiguous pairs -> split arrays, a SIMD de-interleave */
        float* restrict zr = Z->realp; float* restrict zi = Z->imagp; const float* restrict cc = c;
        for (vDSP_Length i = 0; i < N; i++) { zr[i] = cc[2*i]; zi[i] = cc[2*i + 1]; }
        return;
    }
    for (vDSP_Length i = 0; i < N; i++) { Z->realp[i*IZ] = c[i*IC]; Z->imagp[i*IZ] = c[i*IC + 1]; }
}

void vDSP_ztoc(const DSPSplitComplex* Z, vDSP_Stride IZ, DSPComplex* C, vDSP_Stride IC, vDSP_Length N) {
    float* c = (float*)C;
    if (IC == 2 && IZ == 1) {            /* m:355 */
        const float* restrict zr = Z->realp; const float* restrict zi = Z->imagp; float* restrict cc = c;
        for (vDSP_Length i = 0; i < N; i++) { cc[2*i] = zr[i]; cc[2*i + 1] = zi[i]; }
        return;
    }
    for (vDSP_Length i = 0; i < N; i++) { c[i*IC] = Z->realp[i*IZ]; c[i*IC + 1] = Z->imagp[i*IZ]; }
}

/* Stockham autosort radix-2 (decimation in frequency), forward (e^{-i theta}), M complex points.
 * tw holds cos/sin(2*pi*j/N) for j < N/2 with N = 2M, so e^{-2 pi i p/n} is entry p*(N/n).
 * Returns whichever ping-pong buffer holds the result (natural order). */
#define STOCKHAM(T, NAME)                                                                          \
static T* NAME(T* x, T* y, size_t M, const T* tw) {                                                \
    size_t N = 2 * M;                                                                              \
    for (size_t n = M, s = 1; n > 1; n >>= 1, s <<= 1) {                                           \
        size_t m = n / 2, tstep = N / n;                                                           \
        for (size_t p = 0; p < m; p++) {                                                           \
            T wr = tw[2 * p * tstep], wi = -tw[2 * p * tstep + 1];                                  \
            const T* xa = x + 2 * s * p; const T* xb = x + 2 * s * (p + m);                         \
            T* ya = y + 2 * s * (2 * p); T* yb = y + 2 * s * (2 * p + 1);                           \
            for (size_t q = 0; q < s; q++) {                                                       \
                T ar = xa[2*q], ai = xa[2*q+1], br = xb[2*q], bi = xb[2*q+1];                       \
                T dr = ar - br, di = ai - bi;                                                      \
                ya[2*q] = ar + br;            ya[2*q+1] = ai + bi;                                 \
                yb[2*q] = dr * wr - di * wi;  yb[2*q+1] = dr * wi + di * wr;                       \
            }                                                                                      \
        }                                                                                          \
        T* t = x; x = y; y = t;                                                                    \
    }                                                                                              \
    return x;                                                                                      \
}
STOCKHAM(double, stockham_f64)
STOCKHAM(float,  stockham_f32)


/* ---- mode 2: four-step single-precision FFT --------------------------------------------------------------
 * M = N/2 complex points z[n] = x[2n] + i x[2n+1] (what vDSP_ctoz left in realp/imagp), M = 32 R, R = 4 .. 32:
 *   n = R n1 + n2, k = k1 + 32 k2
 *   step 1: 32-point DFT over n1 for every column n2            (rows of R floats: contiguous SIMD loops)
 *   twiddle by exp(-2 pi i n2 k1 / M), transposed to [n2][k1]
 *   step 2: R-point DFT over n2 for every column k1             (rows of 32 floats)
 *   -> Z[k1 + 32 k2] sits at [k2][k1], i.e. in natural order
 *   real split: 2 X[k] = (Z[k] + conj Z[M-k]) - i e^{-2 pi i k / N} (Z[k] - conj Z[M-k])
 * The row transforms are radix-2 decimation in frequency, in place; they leave output index bitrev(r) in row r,
 * which the following step undoes through its row addressing.  Split real / imaginary arrays throughout. */
#if defined(__x86_64__) && defined(__GNUC__)
#define LBAD_CLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))
#else
#define LBAD_CLONES
#endif

struct LBADFastFFT {
    unsigned M, R, logR;
    float *ar, *ai, *br, *bi;          /* [32][R] and [R][32] work arrays, 64-byte aligned */
    float *t1c, *t1s;                  /* step-1 row twiddles: stage s, butterfly j -> cos, sin of 2 pi j 2^s / 32 */
    float *t2c, *t2s;                  /* step-2 row twiddles for R points */
    float *twc, *tws;                  /* inter-step twiddles in the order they are applied: [bitrev5 row r -> k1][n2] */
    float *sc, *ss;                    /* real-split twiddles cos, sin(2 pi k / N), k < M */
    float *zr, *zi;                    /* spectrum of the complex transform, natural order, M + 1 entries (Z[M] = Z[0]) */
};

static unsigned bitrev_n(unsigned v, unsigned bits) { unsigned r = 0; for (unsigned b = 0; b < bits; b++) r |= ((v >> b) & 1u) << (bits - 1 - b); return r; }
static float* alloc64(size_t n) { void* p = NULL; if (posix_memalign(&p, 64, (n * sizeof(float) + 63) & ~(size_t)63) != 0) return NULL; return p; }

static struct LBADFastFFT* fast_new(size_t n) {
    if (n < 256 || n > 2048 || (n & (n - 1))) return NULL;
    struct LBADFastFFT* f = calloc(1, sizeof *f);
    f->M = (unsigned)n / 2; f->R = f->M / 32; f->logR = 0; while ((1u << f->logR) < f->R) f->logR++;
    unsigned M = f->M, R = f->R;
    f->ar = alloc64(M); f->ai = alloc64(M); f->br = alloc64(M); f->bi = alloc64(M);
    f->t1c = alloc64(16 * 5); f->t1s = alloc64(16 * 5); f->t2c = alloc64(16 * 5); f->t2s = alloc64(16 * 5);
    f->twc = alloc64(M); f->tws = alloc64(M); f->sc = alloc64(M); f->ss = alloc64(M); f->zr = alloc64(M + 16); f->zi = alloc64(M + 16);
    for (unsigned st = 0; st < 5; st++) for (unsigned j = 0; j < 16; j++) {
        double a1 = 2.0 * M_PI * (double)((j << st) % 32) / 32.0, a2 = 2.0 * M_PI * (double)((j << st) % R) / (double)R;
        f->t1c[st * 16 + j] = (float)cos(a1); f->t1s[st * 16 + j] = (float)sin(a1);
        f->t2c[st * 16 + j] = (float)cos(a2); f->t2s[st * 16 + j] = (float)sin(a2);
    }
    for (unsigned r = 0; r < 32; r++) for (unsigned n2 = 0; n2 < R; n2++) {
        double a = 2.0 * M_PI * (double)(bitrev_n(r, 5) * n2) / (double)M;
        f->twc[r * R + n2] = (float)cos(a); f->tws[r * R + n2] = (float)sin(a);
    }
    for (unsigned k = 0; k < M; k++) { double a = 2.0 * M_PI * (double)k / (double)n; f->sc[k] = (float)cos(a); f->ss[k] = (float)sin(a); }
    return f;
}
static void fast_free(struct LBADFastFFT* f) {
    if (!f) return;
    free(f->ar); free(f->ai); free(f->br); free(f->bi); free(f->t1c); free(f->t1s); free(f->t2c); free(f->t2s);
    free(f->twc); free(f->tws); free(f->sc); free(f->ss); free(f->zr); free(f->zi); free(f);
}

/* in-place radix-2 DIF over `rows` rows of `cols` floats (forward, e^{-i theta}); row r ends up holding output bitrev(r) */
LBAD_CLONES
static void rows_dif(float* restrict xr, float* restrict xi, unsigned rows, unsigned cols, const float* restrict tc, const float* restrict ts) {
    unsigned st = 0;
    for (unsigned half = rows / 2; half >= 1; half >>= 1, st++) {
        for (unsigned b = 0; b < rows; b += 2 * half) {
            for (unsigned j = 0; j < half; j++) {
                const float c = tc[st * 16 + j], s = ts[st * 16 + j];      /* w = c - i s */
                float* restrict pr = xr + (size_t)(b + j) * cols; float* restrict pi = xi + (size_t)(b + j) * cols;
                float* restrict qr = xr + (size_t)(b + j + half) * cols; float* restrict qi = xi + (size_t)(b + j + half) * cols;
                for (unsigned q = 0; q < cols; q++) {
                    const float ar = pr[q], ai = pi[q], cr = qr[q], ci = qi[q];
                    const float dr = ar - cr, di = ai - ci;
                    pr[q] = ar + cr; pi[q] = ai + ci;
                    qr[q] = dr * c + di * s; qi[q] = di * c - dr * s;
                }
            }
        }
    }
}

LBAD_CLONES
static void fast_zrip(struct LBADFastFFT* f, float* re, float* im) {
    const unsigned M = f->M, R = f->R, logR = f->logR;
    float* restrict ar = f->ar; float* restrict ai = f->ai; float* restrict br = f->br; float* restrict bi = f->bi;
    /* z[R n1 + n2] is already laid out as [n1][n2] */
    memcpy(ar, re, M * sizeof(float)); memcpy(ai, im, M * sizeof(float));
    rows_dif(ar, ai, 32, R, f->t1c, f->t1s);
    /* twiddle (row r holds k1 = bitrev5(r); the table is stored in row order) and transpose to [n2][k1] */
    {
        const float* restrict c = f->twc; const float* restrict s = f->tws;
        for (unsigned i = 0; i < M; i++) { const float xr = ar[i], xi = ai[i]; ar[i] = xr * c[i] + xi * s[i]; ai[i] = xi * c[i] - xr * s[i]; }
    }
    for (unsigned r = 0; r < 32; r++) {
        const unsigned k1 = bitrev_n(r, 5);
        const float* restrict xr = ar + (size_t)r * R; const float* restrict xi = ai + (size_t)r * R;
        for (unsigned n2 = 0; n2 < R; n2++) { br[(size_t)n2 * 32 + k1] = xr[n2]; bi[(size_t)n2 * 32 + k1] = xi[n2]; }
    }
    rows_dif(br, bi, R, 32, f->t2c, f->t2s);
    /* row r of b holds k2 = bitrev(r): natural order Z[k1 + 32 k2] */
    float* restrict zr = f->zr; float* restrict zi = f->zi;
    for (unsigned r = 0; r < R; r++) {
        const unsigned k2 = bitrev_n(r, logR);
        memcpy(zr + (size_t)k2 * 32, br + (size_t)r * 32, 32 * sizeof(float)); memcpy(zi + (size_t)k2 * 32, bi + (size_t)r * 32, 32 * sizeof(float));
    }
    zr[M] = zr[0]; zi[M] = zi[0];
    const float dc = 2.0f * (zr[0] + zi[0]), ny = 2.0f * (zr[0] - zi[0]);
    const float* restrict sc = f->sc; const float* restrict ss = f->ss;
    for (unsigned k = 1; k < M; k++) {
        const float yr = zr[M - k], yi = -zi[M - k];
        const float er = zr[k] + yr, ei = zi[k] + yi, orr = zr[k] - yr, oi = zi[k] - yi;
        const float c = sc[k], sn = ss[k];
        re[k] = er + (c * oi - sn * orr); im[k] = ei - (c * orr + sn * oi);
    }
    re[0] = dc; im[0] = ny;
}

void vDSP_fft_zrip(FFTSetup s, const DSPSplitComplex* C, vDSP_Stride IC, vDSP_Length log2n, FFTDirection dir) {
    if (!s || dir != FFT_FORWARD || IC != 1 || log2n != s->log2n) { fprintf(stderr, "lbad shim: unsupported vDSP_fft_zrip call\n"); abort(); }
    size_t N = s->n, M = N / 2;
    float* re = C->realp; float* im = C->imagp;
    if (lbad_shim_get_fft_mode() == 2 && s->fast) { fast_zrip(s->fast, re, im); return; }
    if (lbad_shim_get_fft_mode() == 0) {
        for (size_t i = 0; i < M; i++) { s->wa[2*i] = re[i]; s->wa[2*i+1] = im[i]; }
        const double* Z = stockham_f64(s->wa, s->wb, M, s->twd);
        /* 2 X[k] = (Z[k] + conj Z[M-k]) - i e^{-2 pi i k/N} (Z[k] - conj Z[M-k]) */
        double dc = 2.0 * (Z[0] + Z[1]), ny = 2.0 * (Z[0] - Z[1]);
        for (size_t k = 1; k < M; k++) {
            double zr = Z[2*k], zi = Z[2*k+1], yr = Z[2*(M-k)], yi = -Z[2*(M-k)+1];
            double er = zr + yr, ei = zi + yi, orr = zr - yr, oi = zi - yi;
            double c = s->twd[2*k], sn = s->twd[2*k+1];      /* w = c - i sn ; -i w = -sn - i c */
            double tr = -sn * orr + c * oi, ti = -c * orr - sn * oi;
            re[k] = (float)(er + tr); im[k] = (float)(ei + ti);
        }
        re[0] = (float)dc; im[0] = (float)ny;
    } else {
        for (size_t i = 0; i < M; i++) { s->fa[2*i] = re[i]; s->fa[2*i+1] = im[i]; }
        const float* Z = stockham_f32(s->fa, s->fb, M, s->twf);
        float dc = 2.0f * (Z[0] + Z[1]), ny = 2.0f * (Z[0] - Z[1]);
        for (size_t k = 1; k < M; k++) {
            float zr = Z[2*k], zi = Z[2*k+1], yr = Z[2*(M-k)], yi = -Z[2*(M-k)+1];
            float er = zr + yr, ei = zi + yi, orr = zr - yr, oi = zi - yi;
            float c = s->twf[2*k], sn = s->twf[2*k+1];
            float tr = -sn * orr + c * oi, ti = -c * orr - sn * oi;
            re[k] = er + tr; im[k] = ei + ti;
        }
        re[0] = dc; im[0] = ny;
    }
}

/* ------------------------------------------------------- ExtAudioFile ---- */

struct LBADShimURL { const float* samples; SInt64 count; };
struct LBADShimExtAudioFile { const float* samples; SInt64 count; SInt64 cursor; };

OSStatus ExtAudioFileOpenURL(CFURLRef inURL, ExtAudioFileRef* outFile) {
    struct LBADShimExtAudioFile* f = calloc(1, sizeof *f);
    f->samples = inURL->samples; f->count = inURL->count; f->cursor = 0;
    *outFile = f;
    return noErr;
}
OSStatus ExtAudioFileDispose(ExtAudioFileRef f) { free(f); return noErr; }
OSStatus ExtAudioFileSetProperty(ExtAudioFileRef f, UInt32 id, UInt32 size, const void* data) {
    (void)f; (void)id; (void)size; (void)data; return noErr;   /* client format is always f32 mono here */
}
OSStatus ExtAudioFileGetProperty(ExtAudioFileRef f, UInt32 id, UInt32* ioSize, void* out) {
    if (id != kExtAudioFileProperty_FileLengthFrames || *ioSize < sizeof(SInt64)) return -50;
    *(SInt64*)out = f->count; return noErr;
}
OSStatus ExtAudioFileRead(ExtAudioFileRef f, UInt32* ioNumberFrames, AudioBufferList* io) {
    SInt64 left = f->count - f->cursor; if (left < 0) left = 0;
    UInt32 n = *ioNumberFrames; if ((SInt64)n > left) n = (UInt32)left;
    memcpy(io->mBuffers[0].mData, f->samples + f->cursor, (size_t)n * sizeof(float));
    f->cursor += n; *ioNumberFrames = n;
    return noErr;
}
OSStatus ExtAudioFileSeek(ExtAudioFileRef f, SInt64 off) { f->cursor = off; return noErr; }

/* ----------------------------------------------------- AudioConverter ---- */

OSStatus AudioConverterNew(const AudioStreamBasicDescription* a, const AudioStreamBasicDescription* b, AudioConverterRef* o) {
    (void)a; (void)b; (void)o; fprintf(stderr, "lbad shim: AudioConverter path reached (dead in the reference)\n"); abort();
}
OSStatus AudioConverterConvertComplexBuffer(AudioConverterRef c, UInt32 n, const AudioBufferList* i, AudioBufferList* o) {
    (void)c; (void)n; (void)i; (void)o; abort();
}
OSStatus AudioConverterDispose(AudioConverterRef c) { (void)c; abort(); }
