/*
 * lbad_math.cuh — per-lane arithmetic of the fingerprint kernels, written as __host__ __device__ inline code so
 * that the exact same functions run inside the CUDA kernels (lbad_extract.cu, lbad_search.cu) and inside the
 * host-side lane emulator used by the CPU unit tests (tests/test_lane_emulation.py -> csrc/lane_emulator.cpp).
 *
 * Windows shorter than 2048 use the same network with several windows per warp: a window of N = 64 R samples (R = 32, 16, 8, 4)
 * is M = 32 R complex points, n = R n1 + n2, k = k1 + 32 k2; 32/R windows sit side by side in the lanes for pass 1 (lane = window
 * * R + n2, the in-lane 32-point DFT over n1 is unchanged), the same 32 x 32 transpose follows, and pass 2 is fft32_tail<R>: one
 * R-point DFT per window in every lane.
 *
 * Layout of the 2048-point real FFT used by the fused kernel (one warp per window):
 *   z[n] = x[2n] + i x[2n+1], n < 1024 = 32 x 32.   n = 32 n1 + n2,  k = k1 + 32 k2.
 *   pass 1: lane n2 holds z[32 n1 + n2] (n1 = register index) and does a 32-point DFT over n1 -> A[n2][k1]
 *   twiddle: A[n2][k1] *= exp(-2 pi i n2 k1 / 1024)
 *   transpose through shared memory so that lane k1 holds A[n2][k1] for all n2
 *   pass 2: 32-point DFT over n2 -> Z[k1 + 32 k2] in lane k1
 *   split : 2 X[k] = (Z[k] + conj Z[1024-k]) - i exp(-2 pi i k / 2048) (Z[k] - conj Z[1024-k])
 * which is vDSP_fft_zrip's output convention (x2, e^{-i theta}) as the reference consumes it (LBAudioDetective.m:353-355).
 * The in-register 32-point DFT is a fully unrolled radix-2 DIF; register position p ends up holding output
 * index bitrev5(p), and callers account for that permutation at compile time.
 */
#ifndef LBAD_MATH_CUH
#define LBAD_MATH_CUH
#include <stdint.h>
#include <math.h>
#include <utility>

#if defined(__CUDACC__)
#define LBAD_HD __host__ __device__ __forceinline__
#else
#define LBAD_HD inline
#endif

namespace lbad {

LBAD_HD constexpr int bitrev5(int p) {
    return ((p & 1) << 4) | ((p & 2) << 2) | (p & 4) | ((p & 8) >> 2) | ((p & 16) >> 4);
}

/* cos / sin of 2*pi*m/32 for m in [0, 16) */
LBAD_HD constexpr float cos32(int m) {
    return m == 0 ? 1.0f : m == 1 ? 0.98078528040323043f : m == 2 ? 0.92387953251128674f : m == 3 ? 0.83146961230254524f
         : m == 4 ? 0.70710678118654752f : m == 5 ? 0.55557023301960218f : m == 6 ? 0.38268343236508977f : m == 7 ? 0.19509032201612825f
         : m == 8 ? 0.0f : -cos32(16 - m);
}
LBAD_HD constexpr float sin32(int m) { return m <= 8 ? cos32(8 - m) : cos32(m - 8); }

#if !defined(__CUDACC__)
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

/* Packed FP32x2 arithmetic: on sm_100a one FADD2 / FFMA2 instruction per (re, im) pair — half the issue slots of the
 * scalar form, same IEEE result per component; on the host (lane emulator) plain component-wise code. */
LBAD_HD float2 add2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
LBAD_HD float2 fma2(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

/* one radix-2 DIF butterfly of a 32-point transform: half-span H, block start S, offset J within the half-block.
 * z[] holds (re, im) pairs; the sum (and the difference when its twiddle is 1) is one packed instruction. */
template <int H, int S, int J>
LBAD_HD void fft32_butterfly(float2 (&z)[32], const float2 neg1) {
    constexpr int a = S + J, b = S + J + H;
    constexpr int m = J * (16 / H);              /* twiddle exp(-2 pi i m / 32), 0 <= m < 16 */
    constexpr float kS2 = 0.70710678118654752f;
    const float2 za = z[a], zb = z[b];
    z[a] = add2(za, zb);
    if constexpr (m == 0) { z[b] = fma2(zb, neg1, za); }                                    /* za - zb, exactly */
    else {
        const float dr = za.x - zb.x, di = za.y - zb.y;
        if constexpr (m == 8)       { z[b] = make_float2(di, -dr); }                                         /* x (-i) */
        else if constexpr (m == 4)  { z[b] = make_float2((dr + di) * kS2, (di - dr) * kS2); }                /* x (1-i)/sqrt2 */
        else if constexpr (m == 12) { z[b] = make_float2((di - dr) * kS2, (dr + di) * (-kS2)); }             /* x (-1-i)/sqrt2 */
        else {
            constexpr float c = cos32(m), s = sin32(m);                                                      /* x (c - i s) */
            z[b] = make_float2(dr * c + di * s, di * c - dr * s);
        }
    }
}

template <int H, int... I>
LBAD_HD void fft32_stage(float2 (&z)[32], const float2 neg1, std::integer_sequence<int, I...>) {
    (fft32_butterfly<H, (I / H) * 2 * H, I % H>(z, neg1), ...);
}

/* The last log2(R) stages of the 32-point DIF network: 32/R independent forward R-point DFTs, one on every aligned block of R
 * register positions (R = 32: the whole 32-point DFT).  Inside a block, position q holds output index bitrev over log2(R) bits. */
template <int R>
LBAD_HD void fft32_tail(float2 (&z)[32]) {
    const float2 neg1 = make_float2(-1.0f, -1.0f);
    using seq = std::make_integer_sequence<int, 16>;
    if constexpr (R >= 32) fft32_stage<16>(z, neg1, seq{});
    if constexpr (R >= 16) fft32_stage<8>(z, neg1, seq{});
    if constexpr (R >= 8) fft32_stage<4>(z, neg1, seq{});
    if constexpr (R >= 4) fft32_stage<2>(z, neg1, seq{});
    if constexpr (R >= 2) fft32_stage<1>(z, neg1, seq{});
}
/* in-place forward 32-point DFT (e^{-i theta}); output index of register position p is bitrev5(p) */
LBAD_HD void fft32(float2 (&z)[32]) { fft32_tail<32>(z); }

LBAD_HD constexpr int log2c(int v) { return v <= 1 ? 0 : 1 + log2c(v >> 1); }
/* bit reversal over log2(R) bits */
template <int R> LBAD_HD constexpr int bitrevR(int q) { return bitrev5(q) >> (5 - log2c(R)); }

/* 2 X[k] from Z[k] = z and Z[M-k] = p, with w = exp(-2 pi i k / N) = (c, -s) given as c, s */
LBAD_HD void real_split_2x(float2 z, float2 p, float c, float s, float& xr, float& xi) {
    const float2 e = fma2(p, make_float2(1.0f, -1.0f), z);     /* Z + conj Z' */
    const float2 d = fma2(p, make_float2(-1.0f, 1.0f), z);     /* Z - conj Z' */
    xr = fmaf(c, d.y, fmaf(-s, d.x, e.x));                     /* -i w d = (-s dr + c di) + i (-c dr - s di) */
    xi = fmaf(-s, d.y, fmaf(-c, d.x, e.y));
}

/* LBAudioDetective.m:387-401 for one bin: positive parts only are divided by pos_scale, then re^2 + im^2; a
 * non-finite value contributes 0 (the reference skips it).  scale_m1 = 1/pos_scale - 1: pos_scale is a power of two,
 * so fma(max(x, 0), scale_m1, x) is x/pos_scale for x > 0 (the exact quotient is representable, one rounding) and x
 * otherwise (NaN propagates and is then dropped, as in the reference where NaN > 0 is false). */
LBAD_HD float bin_energy(float re, float im, float scale_m1) {
    re = fmaf(fmaxf(re, 0.0f), scale_m1, re);
    im = fmaf(fmaxf(im, 0.0f), scale_m1, im);
    const float v = fmaf(re, re, im * im);
    return (v <= 3.402823466e+38f) ? v : 0.0f;
}

/* The same without the finiteness filter, for callers that check the band sum instead: every term is >= 0, so a band sum is finite
 * exactly when all of its terms are, and only a non-finite sum (never seen on real audio) needs the filtered re-summation. */
LBAD_HD float bin_energy_raw(float re, float im, float scale_m1) {
    const float2 g = fma2(make_float2(fmaxf(re, 0.0f), fmaxf(im, 0.0f)), make_float2(scale_m1, scale_m1), make_float2(re, im));
    return fmaf(g.x, g.x, g.y * g.y);
}

/* Hit mask of one 32-pair word (LBAudioDetectiveFingerprint.m:155-169):
 * pairs where fp1 has a bit set (P1|M1) and both bits agree with fp2. */
LBAD_HD uint32_t hit_word(uint32_t p1, uint32_t m1, uint32_t p2, uint32_t m2) {
    return (p1 | m1) & ~(p1 ^ p2) & ~(m1 ^ m2);
}

}  // namespace lbad
#endif
