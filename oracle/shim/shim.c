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
 *       mode 1 "f32"  (used for the CPU timing baseline): a float Stockham radix-2 FFT with table
 *                     twiddles, comparable in cost to a production single-precision FFT.
 *     Select with lbad_shim_set_fft_mode() or the LBAD_SHIM_FFT=f64|f32 environment variable.
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
};

static int g_fft_mode = -1;

void lbad_shim_set_fft_mode(int mode) { g_fft_mode = mode ? 1 : 0; }
int lbad_shim_get_fft_mode(void) {
    if (g_fft_mode < 0) {
        const char* e = getenv("LBAD_SHIM_FFT");
        g_fft_mode = (e && strcmp(e, "f32") == 0) ? 1 : 0;
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
    return s;
}

void vDSP_destroy_fftsetup(FFTSetup s) {
    if (!s) return;   /* the reference destroys a NULL setup on its first SetWindowSize (LBAudioDetective.m:179) */
    free(s->twd); free(s->twf); free(s->wa); free(s->wb); free(s->fa); free(s->fb); free(s);
}

void vDSP_ctoz(const DSPComplex* C, vDSP_Stride IC, const DSPSplitComplex* Z, vDSP_Stride IZ, vDSP_Length N) {
    /* IC is in floats (2 = contiguous complex), as in vDSP */
    const float* c = (const float*)C;
    for (vDSP_Length i = 0; i < N; i++) { Z->realp[i*IZ] = c[i*IC]; Z->imagp[i*IZ] = c[i*IC + 1]; }
}

void vDSP_ztoc(const DSPSplitComplex* Z, vDSP_Stride IZ, DSPComplex* C, vDSP_Stride IC, vDSP_Length N) {
    float* c = (float*)C;
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

void vDSP_fft_zrip(FFTSetup s, const DSPSplitComplex* C, vDSP_Stride IC, vDSP_Length log2n, FFTDirection dir) {
    if (!s || dir != FFT_FORWARD || IC != 1 || log2n != s->log2n) { fprintf(stderr, "lbad shim: unsupported vDSP_fft_zrip call\n"); abort(); }
    size_t N = s->n, M = N / 2;
    float* re = C->realp; float* im = C->imagp;
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
