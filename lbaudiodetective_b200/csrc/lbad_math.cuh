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
 * The in-register 32-point DFT is a fully unrolled radix-2 decimation in time with Linzer-Feig (FMA-fused) butterflies; the input is
 * read in natural order, register position p ends up holding output index bitrev5(p), and callers account for that permutation at
 * compile time.  At the reference's hop of 64 samples pass 1 is further split into two 16-point halves of which one is carried over
 * from the previous window (dft16 / dit32_combine below).
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
LBAD_HD float2 mul2(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
/* The packed instructions take their operands with modifiers — halves swapped, either half negated, one scalar broadcast to both halves
 * (SASS: R.F32x2.LO_HI.NP, R.F32) — so -i b, i b and a broadcast cost nothing, and complex arithmetic written with them is ONE issue slot
 * per (re, im) pair where the scalar form needs two: */
LBAD_HD float2 mul_neg_i(float2 b) { return make_float2(b.y, -b.x); }       /* -i b */
LBAD_HD float2 mul_pos_i(float2 b) { return make_float2(-b.y, b.x); }       /*  i b */
LBAD_HD float2 both(float v) { return make_float2(v, v); }
/* z x t: (z.x t.x - z.y t.y, z.x t.y + z.y t.x) as FMUL2 + FFMA2 */
LBAD_HD float2 cmul(float2 z, float2 t) { return fma2(t, both(z.x), mul2(mul_pos_i(t), both(z.y))); }

/* cos(2 pi m / 32), m = 0..8, to double precision; the remaining angles follow by symmetry */
LBAD_HD constexpr double cos32d(int m) {
    return m == 0 ? 1.0 : m == 1 ? 0.98078528040323044913 : m == 2 ? 0.92387953251128675613 : m == 3 ? 0.83146961230254523708
         : m == 4 ? 0.70710678118654752440 : m == 5 ? 0.55557023301960222474 : m == 6 ? 0.38268343236508977173 : m == 7 ? 0.19509032201612826785
         : m == 8 ? 0.0 : -cos32d(16 - m);
}
LBAD_HD constexpr double sin32d(int m) { return m <= 8 ? cos32d(8 - m) : cos32d(m - 8); }

/* One radix-2 decimation-in-time butterfly, twiddle first: (A, B) <- (A + w B, A - w B), w = exp(-2 pi i M / 32), 0 <= M < 16.
 * z[] holds (re, im) pairs.  The general twiddle is applied in the Linzer-Feig form: with w = c - i s and t = s / c,
 * w B = c (B.x + t B.y, B.y - t B.x), so the product costs two FMAs and its scale c rides on the two packed FMAs that
 * form A +- (...) — six FMA-pipe slots per butterfly instead of eight.  Where |s| > |c| the same is done with s factored
 * out (cotangent form), so the stored ratio never exceeds 1 in magnitude.  The scale is a literal: FFMA2 takes it as a
 * broadcast immediate. */
template <int M>
LBAD_HD void dit_bfly(const float2 A, const float2 B, float2& X, float2& Y) {
    if constexpr (M == 0) {
        X = add2(A, B);
        Y = fma2(B, make_float2(-1.0f, -1.0f), A);
    } else if constexpr (M == 8) {                                  /* w = -i: w B = (B.y, -B.x) */
        X = add2(A, mul_neg_i(B));
        Y = add2(A, mul_pos_i(B));
    } else {
        constexpr double c = cos32d(M), s = sin32d(M);
        constexpr bool tan_form = (c < 0 ? -c : c) >= (s < 0 ? -s : s);
        constexpr float scale = (float)(tan_form ? c : s);
        float2 u;                                                                     /* w B = scale x u, every form one packed instruction */
        if constexpr (M == 4) u = add2(B, mul_neg_i(B));                              /* t = 1:  (B.x + B.y, B.y - B.x) */
        else if constexpr (M == 12) u = add2(B, mul_pos_i(B));                        /* t = -1: (B.x - B.y, B.y + B.x) */
        else if constexpr (tan_form) { constexpr float t = (float)(s / c); u = fma2(mul_neg_i(B), both(t), B); }       /* (B.x + t B.y, B.y - t B.x) */
        else                         { constexpr float t = (float)(c / s); u = fma2(B, both(t), mul_neg_i(B)); }       /* (t B.x + B.y, t B.y - B.x) */
        X = fma2(u, make_float2(scale, scale), A);
        Y = fma2(u, make_float2(-scale, -scale), A);
    }
}
template <int M, int PA, int PB>
LBAD_HD void dit_butterfly(float2 (&z)[32]) { dit_bfly<M>(z[PA], z[PB], z[PA], z[PB]); }

LBAD_HD constexpr int log2c(int v) { return v <= 1 ? 0 : 1 + log2c(v >> 1); }
/* bit reversal over log2(R) bits */
template <int R> LBAD_HD constexpr int bitrevR(int q) { return bitrev5(q) >> (5 - log2c(R)); }

/* stage with logical half-span HALF of the R-point transforms living on the aligned blocks of R register positions:
 * logical index i of a block sits at position bitrevR(i), so the input is read in natural order and output k ends up at bitrevR(k) */
template <int R, int HALF, int... I>
LBAD_HD void dit_stage(float2 (&z)[32], std::integer_sequence<int, I...>) {
    /* butterfly I of the 16 in this stage: block of R positions (I / (R/2)), within it logical pair (a, a + HALF) */
    (dit_butterfly<((I % (R / 2)) % HALF) * (16 / HALF),
                   (I / (R / 2)) * R + bitrevR<R>(((I % (R / 2)) / HALF) * 2 * HALF + (I % (R / 2)) % HALF),
                   (I / (R / 2)) * R + bitrevR<R>(((I % (R / 2)) / HALF) * 2 * HALF + (I % (R / 2)) % HALF + HALF)>(z), ...);
}

/* 32/R independent forward R-point DFTs (e^{-i theta}), one on every aligned block of R register positions (R = 32: one 32-point
 * DFT).  Inside a block the input is in natural order and position q ends up holding output index bitrevR(q). */
template <int R>
LBAD_HD void fft32_tail(float2 (&z)[32]) {
    using seq = std::make_integer_sequence<int, 16>;
    if constexpr (R >= 2) dit_stage<R, 1>(z, seq{});
    if constexpr (R >= 4) dit_stage<R, 2>(z, seq{});
    if constexpr (R >= 8) dit_stage<R, 4>(z, seq{});
    if constexpr (R >= 16) dit_stage<R, 8>(z, seq{});
    if constexpr (R >= 32) dit_stage<R, 16>(z, seq{});
}
/* The 32-point transform of a hop-N/32 sliding window in two halves.  The decimation-in-time split X[k] = E[k] + W^k O[k],
 * X[k+16] = E[k] - W^k O[k] has E = DFT16 of the even-indexed inputs and O = DFT16 of the odd-indexed ones; when consecutive
 * windows are one input apart, the odd inputs of window w ARE the even inputs of window w + 1, so O of one window is carried
 * over as E of the next (same values, same arithmetic — nothing is approximated) and each window costs one 16-point transform
 * plus the combining stage.  dft16: natural order in, output k at position bitrev4(k).  dit32_combine: position q of e / o holds
 * E / O[bitrev4(q)]; output index k1 lands at position bitrev5(k1) of z, the convention of fft32. */
template <int HALF, int... I>
LBAD_HD void dft16_stage(float2 (&h)[16], std::integer_sequence<int, I...>) {
    (dit_bfly<(I % HALF) * (16 / HALF)>(h[bitrevR<16>((I / HALF) * 2 * HALF + I % HALF)], h[bitrevR<16>((I / HALF) * 2 * HALF + I % HALF + HALF)],
                                        h[bitrevR<16>((I / HALF) * 2 * HALF + I % HALF)], h[bitrevR<16>((I / HALF) * 2 * HALF + I % HALF + HALF)]), ...);
}
LBAD_HD void dft16(float2 (&h)[16]) {
    using seq = std::make_integer_sequence<int, 8>;
    dft16_stage<1>(h, seq{}); dft16_stage<2>(h, seq{}); dft16_stage<4>(h, seq{}); dft16_stage<8>(h, seq{});
}
template <int... Q>
LBAD_HD void dit32_combine_impl(const float2 (&e)[16], const float2 (&o)[16], float2 (&z)[32], std::integer_sequence<int, Q...>) {
    (dit_bfly<bitrevR<16>(Q)>(e[Q], o[Q], z[2 * Q], z[2 * Q + 1]), ...);
}
LBAD_HD void dit32_combine(const float2 (&e)[16], const float2 (&o)[16], float2 (&z)[32]) {
    dit32_combine_impl(e, o, z, std::make_integer_sequence<int, 16>{});
}

/* The same transforms with the input given as separate real and imaginary arrays (what the component-wise shared-memory
 * transposition delivers, four consecutive positions per 128-bit load): the first stage, whose twiddles are all 1, is done in
 * scalar adds that write the (re, im) register pairs directly, so no moves are needed to assemble pairs for the packed stages. */
template <int R, int... I>
LBAD_HD void dit_first_stage_soa(const float (&zx)[32], const float (&zy)[32], float2 (&z)[32], std::integer_sequence<int, I...>) {
    ((z[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)))] =
          make_float2(zx[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)))] + zx[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)) + 1)],
                      zy[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)))] + zy[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)) + 1)]),
      z[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)) + 1)] =
          make_float2(zx[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)))] - zx[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)) + 1)],
                      zy[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)))] - zy[(I / (R / 2)) * R + bitrevR<R>(2 * (I % (R / 2)) + 1)])), ...);
}
template <int R>
LBAD_HD void fft32_tail_soa(const float (&zx)[32], const float (&zy)[32], float2 (&z)[32]) {
    using seq = std::make_integer_sequence<int, 16>;
    static_assert(R >= 2, "at least one stage");
    dit_first_stage_soa<R>(zx, zy, z, seq{});
    if constexpr (R >= 4) dit_stage<R, 2>(z, seq{});
    if constexpr (R >= 8) dit_stage<R, 4>(z, seq{});
    if constexpr (R >= 16) dit_stage<R, 8>(z, seq{});
    if constexpr (R >= 32) dit_stage<R, 16>(z, seq{});
}
/* in-place forward 32-point DFT (e^{-i theta}); output index of register position p is bitrev5(p) */
LBAD_HD void fft32(float2 (&z)[32]) { fft32_tail<32>(z); }


/* 2 X[k] from Z[k] = z and Z[M-k] = p, with w = exp(-2 pi i k / N) = (c, -s) given as c, s */
LBAD_HD void real_split_2x(float2 z, float2 p, float c, float s, float& xr, float& xi) {
    const float2 e = fma2(p, make_float2(1.0f, -1.0f), z);     /* Z + conj Z' */
    const float2 d = fma2(p, make_float2(-1.0f, 1.0f), z);     /* Z - conj Z' */
    /* -i w d = (-s dr + c di) + i (-c dr - s di): (xr, xi) = e + (-s, -c) dr + (c, -s) di, two packed FMAs */
    const float2 x = fma2(make_float2(c, -s), both(d.y), fma2(make_float2(-s, -c), both(d.x), e));
    xr = x.x; xi = x.y;
}

/* Both members of a mirrored pair of bins from one evaluation: with Z[k] = z and Z[M-k] = p (w as above),
 * lo = 2 X[k] and hi = conj(2 X[M-k]): the sum e = z + conj p and the rotated difference g = -i w (z - conj p) are shared,
 * 2 X[k] = e + g and 2 X[M-k] = conj(e - g).  g is kept as (Re g, -Im g) so that both results are one packed FMA. */
LBAD_HD void real_split_pair_2x(float2 z, float2 p, float c, float s, float2& lo, float2& hi) {
    const float2 e = fma2(p, make_float2(1.0f, -1.0f), z);
    const float2 d = fma2(p, make_float2(-1.0f, 1.0f), z);
    const float2 g = fma2(make_float2(c, s), both(d.y), mul2(make_float2(-s, c), both(d.x)));     /* (c di - s dr, s di + c dr) */
    lo = fma2(g, make_float2(1.0f, -1.0f), e);
    hi = fma2(g, make_float2(-1.0f, 1.0f), e);
}

/* cos(pi m / 32), m = 0..16, to double precision (the angle of the real-split twiddle advances by pi/32 per row of 32 bins) */
LBAD_HD constexpr double cos64d(int m) {
    return m == 0 ? 1.0 : m == 1 ? 0.99518472667219688624 : m == 2 ? 0.98078528040323044913 : m == 3 ? 0.95694033573220886494
         : m == 4 ? 0.92387953251128675613 : m == 5 ? 0.88192126434835502971 : m == 6 ? 0.83146961230254523708 : m == 7 ? 0.77301045336273696081
         : m == 8 ? 0.70710678118654752440 : m == 9 ? 0.63439328416364549822 : m == 10 ? 0.55557023301960222474 : m == 11 ? 0.47139673682599764856
         : m == 12 ? 0.38268343236508977173 : m == 13 ? 0.29028467725446236764 : m == 14 ? 0.19509032201612826785 : m == 15 ? 0.09801714032956060199 : 0.0;
}
LBAD_HD constexpr double sin64d(int m) { return cos64d(16 - m); }

/* The real split of row K2 (bins k = lane + 32 K2) without a twiddle table: exp(i 2 pi k / N) = exp(i 2 pi lane / N) x exp(i pi K2 / 32), a
 * per-lane factor wl = (cos, sin) times a compile-time one.  With D = d.y + i d.x (d = Z[k] - conj Z[M-k]) the rotated difference of
 * real_split_pair_2x is G = exp(i theta) D.  The row's rotation goes first, in the Linzer-Feig form — ONE packed FMA, its scale rides on
 * the packed FMAs that form e +- G — then the lane's (two packed instructions, as with a table): the same instruction count as the table
 * form, one table load (two shared-memory wavefronts) less per row. */
template <int K2>
LBAD_HD void real_split_pair_row(float2 z, float2 p, float2 wl, float2& lo, float2& hi) {
    static_assert(K2 >= 0 && K2 <= 16, "rows of the lower half spectrum");
    const float2 e = fma2(p, make_float2(1.0f, -1.0f), z);
    const float2 d = fma2(p, make_float2(-1.0f, 1.0f), z);
    constexpr double c = cos64d(K2), s = sin64d(K2);
    constexpr float scale = (float)(c >= s ? c : s);
    const float2 D = make_float2(d.y, d.x), iD = make_float2(-d.x, d.y);             /* D and i D: operand modifiers, no instructions */
    float2 U;                                                                        /* exp(i theta_row) D = scale x U */
    if constexpr (K2 == 0) U = D;
    else if constexpr (c >= s) { constexpr float t = (float)(s / c); U = fma2(iD, both(t), D); }      /* (1 + i t) D */
    else { constexpr float t = (float)(c / s); U = fma2(D, both(t), iD); }                             /* (t + i) D */
    const float2 g = cmul(U, wl);
    lo = fma2(g, make_float2(scale, -scale), e);
    hi = fma2(g, make_float2(-scale, scale), e);
}
/* the same with the row given as a value: inside a fully unrolled loop the switch folds to the one case */
LBAD_HD void real_split_pair_rows(const int k2, float2 z, float2 p, float2 wl, float2& lo, float2& hi) {
    switch (k2) {
        case 0: real_split_pair_row<0>(z, p, wl, lo, hi); break;   case 1: real_split_pair_row<1>(z, p, wl, lo, hi); break;
        case 2: real_split_pair_row<2>(z, p, wl, lo, hi); break;   case 3: real_split_pair_row<3>(z, p, wl, lo, hi); break;
        case 4: real_split_pair_row<4>(z, p, wl, lo, hi); break;   case 5: real_split_pair_row<5>(z, p, wl, lo, hi); break;
        case 6: real_split_pair_row<6>(z, p, wl, lo, hi); break;   case 7: real_split_pair_row<7>(z, p, wl, lo, hi); break;
        case 8: real_split_pair_row<8>(z, p, wl, lo, hi); break;   case 9: real_split_pair_row<9>(z, p, wl, lo, hi); break;
        case 10: real_split_pair_row<10>(z, p, wl, lo, hi); break; case 11: real_split_pair_row<11>(z, p, wl, lo, hi); break;
        case 12: real_split_pair_row<12>(z, p, wl, lo, hi); break; case 13: real_split_pair_row<13>(z, p, wl, lo, hi); break;
        case 14: real_split_pair_row<14>(z, p, wl, lo, hi); break; default: real_split_pair_row<15>(z, p, wl, lo, hi); break;
    }
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

/* bin_energy_raw of the conjugate: the energy of (re, -im) given (re, im) — the positive-part scaling applies to -im, i.e. to
 * the negative part of im */
LBAD_HD float bin_energy_raw_conj(float re, float im, float scale_m1) {
    const float2 g = fma2(make_float2(fmaxf(re, 0.0f), fminf(im, 0.0f)), make_float2(scale_m1, scale_m1), make_float2(re, im));
    return fmaf(g.x, g.x, g.y * g.y);
}

/* Hit mask of one 32-pair word (LBAudioDetectiveFingerprint.m:155-169):
 * pairs where fp1 has a bit set (P1|M1) and both bits agree with fp2. */
LBAD_HD uint32_t hit_word(uint32_t p1, uint32_t m1, uint32_t p2, uint32_t m2) {
    return (p1 | m1) & ~(p1 ^ p2) & ~(m1 ^ m2);
}

}  // namespace lbad
#endif
