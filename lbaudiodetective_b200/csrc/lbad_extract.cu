/*
 * lbad_extract.cu — fingerprint extraction kernels for sm_100a.
 *
 * Replaces, on the GPU, the reference's whole extraction path (file:line into /root/reference/LBAudioDetective/):
 *   framing loop                      LBAudioDetective.m:250-293
 *   vDSP real FFT                     LBAudioDetective.m:351-355
 *   band energies (Q4/Q5/Q6 quirks)   LBAudioDetective.m:373-405
 *   2-D standard Haar                 LBAudioDetectiveFrame.m:113-153
 *   ordered top-t sign bits           LBAudioDetectiveFrame.m:165-191, truncation LBAudioDetective.m:321-328
 *
 * Two paths, each two kernels (spectral images in between, 16 KB per subfingerprint, L2-friendly):
 *   fast    (window 2048/1024/512/256, 32 bands, even hop, frame span fits shared memory):
 *           bands_fused_kernel — persistent CTAs, one frame (= one subfingerprint) per CTA iteration.  The frame's 127*hop+window
 *           samples are brought into shared memory once by a 1-D TMA bulk copy (cp.async.bulk + mbarrier), so every PCM sample is
 *           read from HBM once; each warp runs the real FFT of a window entirely in registers (2 x radix-32 decimation in time with
 *           one shared-memory transpose; at hop 64 the first pass is shared between consecutive windows) and bins the bands.
 *           haar_select32_kernel — one CTA per image: warp-local Haar, radix selection of the t-th magnitude, ranking, packing.
 *   generic (any supported geometry): bands_generic_kernel (one warp per window, shared-memory radix-2 FFT) and
 *           haar_select_kernel<B>.  Also the on-device cross-check of the fast path (the two FFTs are independent implementations).
 */
#include "lbad_common.cuh"
#include "lbad_math.cuh"
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <mutex>
#include <utility>

namespace lbad {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    fprintf(stderr, "LBAudioDetective(CUDA): %s\n", g_err);
}

/* ------------------------------------------------------------------------------------------------ tables ---- */

struct BandTable {                       /* by-value kernel parameter */
    uint32_t klow[LBAD_MAX_BANDS];
    uint32_t khigh[LBAD_MAX_BANDS];
    uint32_t split[LBAD_MAX_BANDS];      /* register-FFT kernel: the two lanes of band b sum [klow, split) and [split, khigh) */
    float    divisor[LBAD_MAX_BANDS];
};

struct Geo {                             /* by-value kernel parameter */
    uint32_t window, log2m, stride, bands, pairs, words_per_plane;
    uint32_t kmin, kmax;                 /* union of the band ranges: [kmin, kmax) */
    float inv_pos_scale;
    uint32_t frames_per_clip;
    uint64_t clip_stride;
};

/* --------------------------------------------------------------------------- Haar + select building blocks ---- */

struct SelectSmem {
    uint32_t counters[32];
    uint32_t warp_tot[32];
    uint32_t surv_key[256];
    uint32_t surv_idx[256];              /* flat index | sign code << 16 (1: v > 0, 2: v < 0) */
    uint32_t words[16];
    uint32_t nsurv, nbucket, threshold;
    uint32_t pad[1];
};

/* Ordered top-T by |v| (stable: lower flat index first on ties) -> packed sign planes.
 * THREADS threads, each owning E consecutive flat indices [tid*E, tid*E+E); elem(idx) returns the coefficient.
 * Must be called by all threads of the CTA.  out: 2*W words (global).  bucket: shared scratch for THREADS*E keys.
 *
 * The T-th largest key (|v| as an integer) is found by bisection on its bits: the 8 exponent bits by every thread
 * over its own keys, then — since only the keys sharing that exponent are still undecided — the 23 mantissa bits by
 * warp 0 alone over the compacted bucket, without block-wide barriers. */
template <int THREADS, int E, class Elem>
__device__ __forceinline__ void select_and_pack(Elem elem, const int T, const int W, uint32_t* __restrict__ out, SelectSmem& sm, uint32_t* bucket) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t key[E];
    uint32_t pos = 0, neg = 0;           /* sign bits of my E elements (E <= 32) */
#pragma unroll
    for (int e = 0; e < E; e++) {
        const float v = elem(tid * E + e);
        key[e] = __float_as_uint(v) & 0x7fffffffu;
        pos |= (v > 0.0f ? 1u : 0u) << e;
        neg |= (v < 0.0f ? 1u : 0u) << e;
    }
    if (tid < 32) sm.counters[tid] = 0;
    if (tid < 16) sm.words[tid] = 0;
    if (tid == 0) { sm.nsurv = 0; sm.nbucket = 0; }
    __syncthreads();
    /* largest exponent-aligned threshold with count(key >= threshold) >= T */
    uint32_t thr = 0;
    for (int bit = 30; bit >= 23; --bit) {
        const uint32_t cand = thr | (1u << bit);
        uint32_t c = 0;
#pragma unroll
        for (int e = 0; e < E; e++) c += (key[e] >= cand) ? 1u : 0u;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(&sm.counters[bit], c);
        __syncthreads();
        if (sm.counters[bit] >= (uint32_t)T) thr = cand;
    }
    /* keys in [thr, thr + 2^23) are undecided; keys above are certainly selected */
    {
        const uint32_t hi = thr + (1u << 23);
        uint32_t above = 0, mine = 0;
#pragma unroll
        for (int e = 0; e < E; e++) { above += key[e] >= hi; mine += (key[e] >= thr && key[e] < hi); }
        const uint32_t wa = __reduce_add_sync(0xffffffffu, above);
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        uint32_t base = 0;
        if (lane == 31) { base = atomicAdd(&sm.nbucket, incl); if (wa) atomicAdd(&sm.counters[22], wa); }
        base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
#pragma unroll
        for (int e = 0; e < E; e++) if (key[e] >= thr && key[e] < hi) bucket[base++] = key[e];
    }
    __syncthreads();
    if (wid == 0) {
        const uint32_t nb = sm.nbucket, need = (uint32_t)T - sm.counters[22];      /* rank of the threshold inside the bucket, >= 1 */
        uint32_t t2 = thr;
        for (int bit = 22; bit >= 0; --bit) {
            const uint32_t cand = t2 | (1u << bit);
            uint32_t c = 0;
            for (uint32_t i = lane; i < nb; i += 32) c += (bucket[i] >= cand) ? 1u : 0u;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= need) t2 = cand;
        }
        if (lane == 0) sm.threshold = t2;
    }
    __syncthreads();
    thr = sm.threshold;
    /* strictly greater: all survive; equal: the first (T - n_gt) in flat-index order survive */
    uint32_t ngt = 0, neq = 0;
#pragma unroll
    for (int e = 0; e < E; e++) { ngt += key[e] > thr; neq += key[e] == thr; }
    uint32_t packed = (ngt << 16) | neq;                 /* both totals <= 8192 */
    uint32_t incl = packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) sm.warp_tot[wid] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) { const uint32_t t = sm.warp_tot[w]; total += t; if (w < wid) before += t; }
    const uint32_t eq_before = ((before + incl - packed) & 0xffffu);
    const uint32_t need_eq = (uint32_t)T - (total >> 16);
    uint32_t eq_rank = eq_before;
#pragma unroll
    for (int e = 0; e < E; e++) {
        bool take = key[e] > thr;
        if (key[e] == thr) { take = eq_rank < need_eq; eq_rank++; }
        if (take) {
            const uint32_t slot = atomicAdd(&sm.nsurv, 1u);
            if (slot < 256u) {
                sm.surv_key[slot] = key[e];
                sm.surv_idx[slot] = (uint32_t)(tid * E + e) | (((pos >> e) & 1u) << 16) | (((neg >> e) & 1u) << 17);
            }
        }
    }
    __syncthreads();
    /* rank by counting: survivor j precedes s iff key_j > key_s, or equal keys and idx_j < idx_s */
    for (int s = tid; s < T; s += THREADS) {
        const uint32_t ks = sm.surv_key[s], is = sm.surv_idx[s], idx = is & 0xffffu;
        uint32_t rank = 0;
        for (int j = 0; j < T; j++) {
            const uint32_t kj = sm.surv_key[j], ij = sm.surv_idx[j] & 0xffffu;
            rank += (kj > ks || (kj == ks && ij < idx)) ? 1u : 0u;
        }
        if (is & (1u << 16)) atomicOr(&sm.words[rank >> 5], 1u << (rank & 31));
        if (is & (1u << 17)) atomicOr(&sm.words[W + (rank >> 5)], 1u << (rank & 31));
    }
    __syncthreads();
    if (tid < 2 * W) out[tid] = sm.words[tid];
}

/* 1-D Haar of LBAudioDetectiveFrame.m:134-153 on a strided vector, sequential, any length (integer halving). */
__device__ __forceinline__ void haar_1d_seq(float* a, int stride, int n, float* tmp, int tstride) {
    const float sn = sqrtf((float)n), s2 = sqrtf(2.0f);
    for (int i = 0; i < n; i++) a[i * stride] = __fdiv_rn(a[i * stride], sn);
    while (n > 1) {
        n /= 2;
        for (int i = 0; i < n; i++) {
            const float x0 = a[(2 * i) * stride], x1 = a[(2 * i + 1) * stride];
            tmp[i * tstride] = __fdiv_rn(__fadd_rn(x0, x1), s2);
            tmp[(n + i) * tstride] = __fdiv_rn(__fsub_rn(x0, x1), s2);
        }
        for (int i = 0; i < 2 * n; i++) a[i * stride] = tmp[i * tstride];
    }
}

/* ------------------------------------------------------------------------------------------ generic path ---- */

constexpr int GEN_WARPS = 4;

/* One warp per window: shared-memory radix-2 DIF FFT of M = N/2 complex points, real split on the band bins only,
 * band sums in the reference's order.  Writes images[frame][row][band]. */
__global__ void __launch_bounds__(GEN_WARPS * 32)
bands_generic_kernel(const float* __restrict__ pcm, float* __restrict__ images, const float2* __restrict__ tw_m,
                     const float2* __restrict__ tw_n, const Geo g, const BandTable bt, const uint64_t total_windows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t M = g.window / 2;
    float2* z = reinterpret_cast<float2*>(smem_raw) + (size_t)wid * M;
    float* v = reinterpret_cast<float*>(reinterpret_cast<float2*>(smem_raw) + (size_t)GEN_WARPS * M) + (size_t)wid * M;
    const uint64_t wpc = (uint64_t)g.frames_per_clip * LBAD_ROWS_PER_FRAME;      /* windows used per clip */
    for (uint64_t w = (uint64_t)blockIdx.x * GEN_WARPS + wid; w < total_windows; w += (uint64_t)gridDim.x * GEN_WARPS) {
        const uint64_t clip = w / wpc, wi = w % wpc;
        const float* x = pcm + clip * g.clip_stride + wi * g.stride;          /* m:262-290: window i starts at hop*i */
        for (uint32_t i = lane; i < M; i += 32) z[i] = make_float2(x[2 * i], x[2 * i + 1]);   /* vDSP_ctoz, m:353 */
        __syncwarp();
        for (uint32_t h = M / 2; h >= 1; h >>= 1) {
            const uint32_t tstep = (M / 2) / h;
            for (uint32_t t = lane; t < M / 2; t += 32) {
                const uint32_t j = t & (h - 1), a = ((t / h) * 2 * h) + j, b = a + h;
                const float2 za = z[a], zb = z[b], tw = tw_m[j * tstep];      /* tw = (cos, -sin) */
                const float dr = za.x - zb.x, di = za.y - zb.y;
                z[a] = make_float2(za.x + zb.x, za.y + zb.y);
                z[b] = make_float2(dr * tw.x - di * tw.y, dr * tw.y + di * tw.x);
            }
            __syncwarp();
        }
        const uint32_t sh = 32 - g.log2m;
        for (uint32_t k = g.kmin + lane; k < g.kmax; k += 32) {
            float xr, xi;
            if (k == 0) {                                                       /* DC in re, Nyquist packed in im (Q2) */
                const float2 z0 = z[0];
                xr = 2.0f * (z0.x + z0.y); xi = 2.0f * (z0.x - z0.y);
            } else {
                const float2 zk = z[__brev(k) >> sh], zp = z[__brev(M - k) >> sh], tw = tw_n[k];   /* tw = (cos, sin) */
                real_split_2x(zk, zp, tw.x, tw.y, xr, xi);
            }
            v[k] = bin_energy(xr, xi, g.inv_pos_scale - 1.0f);                         /* m:387-401 */
        }
        __syncwarp();
        for (uint32_t b = lane; b < g.bands; b += 32) {                         /* m:379-405, sums in increasing k */
            float p = 0.0f;
            for (uint32_t k = bt.klow[b]; k < bt.khigh[b]; k++) p = __fadd_rn(p, v[k]);
            images[w * g.bands + b] = __fdiv_rn(p, bt.divisor[b]);
        }
        __syncwarp();
    }
}

constexpr int HS_THREADS = 256;

/* One CTA per frame: Haar (rows then columns, Frame.m:113-132) + ordered top-T + pack. */
template <int B>
__global__ void __launch_bounds__(HS_THREADS)
haar_select_kernel(const float* __restrict__ images, float* __restrict__ haar_out, uint32_t* __restrict__ words,
                   const int T, const int W, const uint32_t total_frames) {
    constexpr int R = LBAD_ROWS_PER_FRAME, LD = B + 1;
    extern __shared__ __align__(16) float hs_smem[];
    float* a = hs_smem;
    float* tmp = hs_smem + R * LD;
    __shared__ SelectSmem sel;
    const int tid = threadIdx.x;
    for (uint32_t f = blockIdx.x; f < total_frames; f += gridDim.x) {
        const float* img = images + (size_t)f * R * B;
        for (int i = tid; i < R * B; i += HS_THREADS) a[(i / B) * LD + (i % B)] = img[i];
        __syncthreads();
        if (tid < R) haar_1d_seq(a + tid * LD, 1, B, tmp + tid * LD, 1);                 /* rows, Frame.m:114-116 */
        __syncthreads();
        if (tid < B) haar_1d_seq(a + tid, LD, R, tmp + tid, LD);                          /* columns, Frame.m:118-131 */
        __syncthreads();
        if (haar_out) for (int i = tid; i < R * B; i += HS_THREADS) haar_out[(size_t)f * R * B + i] = a[(i / B) * LD + (i % B)];
        select_and_pack<HS_THREADS, (R * B) / HS_THREADS>([&](int idx) { return a[(idx / B) * LD + (idx % B)]; }, T, W,
                                                          words + (size_t)f * 2 * W, sel, reinterpret_cast<uint32_t*>(tmp));
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------------------- Haar + select, 128 x 32 ---- */

/* One CTA per spectral image (128 x 32): Haar rows + columns (Frame.m:113-153), ordered top-T and packing (Frame.m:165-191).
 * The Haar transform is warp-local — every level runs in registers or through warp shuffles — so the whole image needs ONE block
 * barrier (between the row and the column pass); the coefficients never go back to shared memory: each thread keeps its 16 and
 * the selection works on registers.  About 27 KB of shared memory and 64 registers: four CTAs per SM hide each other's barriers. */
constexpr int HS32_THREADS = 256;
constexpr int HS32_LDT = 132;                /* column-major image, imgT[col * 132 + row]: LDS.128-aligned and conflict-free */

/* x / c for a compile-time constant c with r = RN(1/c): multiply + two FMAs give the IEEE quotient for every finite x with
 * |x| >= 2^-100 or x == 0 (exhaustively checked on the host for c = sqrtf(2), sqrtf(32), sqrtf(128)); the rare rest divides. */
constexpr uint32_t DIVC_LO = 0x0d802f51u, DIVC_HI_IMAGE = 0x7c70bdc2u;    /* bit patterns of 7.9e-31f and 1.0e37f / 2 */
__device__ __forceinline__ float div_const(const float x, const float c, const float r) {
    const float q0 = __fmul_rn(x, r);
    const float q = fmaf(fmaf(-q0, c, x), r, q0);
    const float ax = fabsf(x);
    return ((ax >= 7.9e-31f && ax <= 1.0e37f) || ax == 0.0f) ? q : __fdiv_rn(x, c);
}
/* The range of the dividends seen so far, kept as integers: lo = min(|x| bits - 1) over every dividend (so that zero never counts
 * as small); hi = max(|x| bits) over the IMAGE VALUES only — every dividend of the transform is a sum or difference of two values
 * that are themselves bounded by the largest image value M (each level divides by sqrt 2 what the previous one at most doubled, and
 * the leading divisions by sqrt 32 / sqrt 128 undo the five / seven levels), so |dividend| <= 2 M and M <= 1e37 / 2 is enough.
 * FAST mode divides by the short form unconditionally and only records the range; the caller checks it once per image (block-wide)
 * and redoes the image with the checked form if anything fell outside — which real spectra never do. */
struct DivRange {
    uint32_t lo = 0xffffffffu, hi = 0u;
    __device__ __forceinline__ void image_value(const float x) { hi = max(hi, __float_as_uint(x) & 0x7fffffffu); }
    __device__ __forceinline__ bool bad() const { return lo < DIVC_LO - 1u || hi > DIVC_HI_IMAGE; }
};
template <bool FAST>
__device__ __forceinline__ float div_c(const float x, const float c, const float r, DivRange& rg) {
    if constexpr (!FAST) return div_const(x, c, r);
    else {
        const float q0 = __fmul_rn(x, r);
        const uint32_t u = __float_as_uint(x) & 0x7fffffffu;
        rg.lo = min(rg.lo, u - 1u);
        return fmaf(fmaf(-q0, c, x), r, q0);
    }
}

/* four levels of the ordered Haar pyramid (Frame.m:143-152) on 16 consecutive elements held in registers:
 * d1[i] = level-1 difference of pair i (8), d2 (4), d3 (2), d4 (1) and the remaining sum s4 */
template <bool FAST>
__device__ __forceinline__ void haar16(const float (&x)[16], float (&d1)[8], float (&d2)[4], float (&d3)[2], float& d4, float& s4,
                                       const float s2, const float r2, DivRange& rg) {
    float s1[8], t2[4], t3[2];
#pragma unroll
    for (int i = 0; i < 8; i++) { s1[i] = div_c<FAST>(__fadd_rn(x[2 * i], x[2 * i + 1]), s2, r2, rg); d1[i] = div_c<FAST>(__fsub_rn(x[2 * i], x[2 * i + 1]), s2, r2, rg); }
#pragma unroll
    for (int i = 0; i < 4; i++) { t2[i] = div_c<FAST>(__fadd_rn(s1[2 * i], s1[2 * i + 1]), s2, r2, rg); d2[i] = div_c<FAST>(__fsub_rn(s1[2 * i], s1[2 * i + 1]), s2, r2, rg); }
#pragma unroll
    for (int i = 0; i < 2; i++) { t3[i] = div_c<FAST>(__fadd_rn(t2[2 * i], t2[2 * i + 1]), s2, r2, rg); d3[i] = div_c<FAST>(__fsub_rn(t2[2 * i], t2[2 * i + 1]), s2, r2, rg); }
    s4 = div_c<FAST>(__fadd_rn(t3[0], t3[1]), s2, r2, rg);
    d4 = div_c<FAST>(__fsub_rn(t3[0], t3[1]), s2, r2, rg);
}

/* Suffix search in a 256-bin histogram, done redundantly by every warp (no broadcast, no barrier): bin = the largest b with
 * S[b] = sum_{j >= b} hist[j] >= need, above = S[b + 1], count = hist[b].  Requires 1 <= need <= S[0]. */
__device__ __forceinline__ void warp_find_bin256(const uint32_t* __restrict__ hist, const uint32_t need, uint32_t& bin, uint32_t& above, uint32_t& count) {
    const int lane = threadIdx.x & 31;
    const uint4 a = *reinterpret_cast<const uint4*>(hist + 8 * lane), b = *reinterpret_cast<const uint4*>(hist + 8 * lane + 4);
    const uint32_t c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const uint32_t tot = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
    uint32_t suf = tot;                                      /* inclusive suffix over the lanes */
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_down_sync(0xffffffffu, suf, d); if (lane + d < 32) suf += o; }
    uint32_t run = suf - tot, mybin = 0, myabove = 0, mycount = 0;
    const bool here = suf >= need && run < need;             /* true in exactly one lane */
#pragma unroll
    for (int j = 7; j >= 0; --j) {
        if (run < need && run + c[j] >= need) { mybin = 8u * lane + j; myabove = run; mycount = c[j]; }
        run += c[j];
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, here)) - 1;
    bin = __shfl_sync(0xffffffffu, mybin, src); above = __shfl_sync(0xffffffffu, myabove, src); count = __shfl_sync(0xffffffffu, mycount, src);
}

struct Select32Smem {
    uint32_t hist[HS32_THREADS / 32][256];   /* per-warp histograms of the 8 exponent bits */
    __align__(16) uint32_t digit_hist[3][256];   /* mantissa bits 22..15, 14..7, 6..0 of the keys that are still candidates for the threshold */
    uint32_t warp_tot[HS32_THREADS / 32];
    __align__(16) uint32_t surv_key[256];
    uint32_t surv_idx[256];                  /* flat index | sign code << 16 (bit 16: v > 0, bit 17: v < 0) */
    uint32_t words[16];
    uint32_t nsurv, nbucket, expo, above, rank_sum;
    uint32_t steps[16];
};

/* Rows (length 32, Frame.m:114-116) of one 128 x 32 image into the column-major tile imgT, with the column pass's leading division
 * by sqrtf(128) (Frame.m:137-139) folded into the store: warp w owns rows 16w..16w+15, two lanes per row, 16 elements each. */
template <bool FAST, int SRC_LD>     /* SRC_LD = 0: img is in global memory (rows of 32 floats); otherwise a shared-memory image with rows SRC_LD floats apart */
__device__ __forceinline__ void haar32_rows(const float* __restrict__ img, float* __restrict__ imgT, DivRange& rg) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float s2 = sqrtf(2.0f), s32 = sqrtf(32.0f), s128 = sqrtf(128.0f);
    const float r2 = 1.0f / s2, r32 = 1.0f / s32, r128 = 1.0f / s128;
    const int row = 16 * wid + (lane >> 1), half = lane & 1;
    const float4* src = reinterpret_cast<const float4*>(img + (size_t)row * (SRC_LD ? SRC_LD : 32) + half * 16);
    float x[16], d1[8], d2[4], d3[2], d4, s4;
#pragma unroll
    for (int j = 0; j < 4; j++) { const float4 v = SRC_LD ? src[j] : __ldg(src + j); x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w; }
#pragma unroll
    for (int j = 0; j < 16; j++) { if constexpr (FAST) rg.image_value(x[j]); x[j] = div_c<FAST>(x[j], s32, r32, rg); }   /* Frame.m:137-139 */
    haar16<FAST>(x, d1, d2, d3, d4, s4, s2, r2, rg);
    const float other = __shfl_xor_sync(0xffffffffu, s4, 1);                                /* level 5 joins the two halves */
    const float top = half ? div_c<FAST>(__fsub_rn(other, s4), s2, r2, rg) : div_c<FAST>(__fadd_rn(s4, other), s2, r2, rg);
    float* dst = imgT + row;                                                                /* ordered output positions */
#pragma unroll
    for (int i = 0; i < 8; i++) dst[(16 + 8 * half + i) * HS32_LDT] = div_c<FAST>(d1[i], s128, r128, rg);
#pragma unroll
    for (int i = 0; i < 4; i++) dst[(8 + 4 * half + i) * HS32_LDT] = div_c<FAST>(d2[i], s128, r128, rg);
#pragma unroll
    for (int i = 0; i < 2; i++) dst[(4 + 2 * half + i) * HS32_LDT] = div_c<FAST>(d3[i], s128, r128, rg);
    dst[(2 + half) * HS32_LDT] = div_c<FAST>(d4, s128, r128, rg);
    dst[half * HS32_LDT] = div_c<FAST>(top, s128, r128, rg);
}

/* Columns (length 128, Frame.m:118-131): warp w owns columns 4w..4w+3, eight lanes per column, 16 rows each; four levels in
 * registers and three by shuffles.  The thread keeps its 16 coefficients (coef), whose flat indices are given by haar32_flat_idx. */
template <bool FAST>
__device__ __forceinline__ void haar32_cols(const float* __restrict__ imgT, float (&coef)[16], DivRange& rg) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float s2 = sqrtf(2.0f), r2 = 1.0f / s2;
    const int cl = lane & 3, g = lane >> 2, col = 4 * wid + cl;
    float x[16], d1[8], d2[4], d3[2], d4, s4;
    const float4* src = reinterpret_cast<const float4*>(imgT + col * HS32_LDT + 16 * g);
#pragma unroll
    for (int j = 0; j < 4; j++) { const float4 v = src[j]; x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w; }
    haar16<FAST>(x, d1, d2, d3, d4, s4, s2, r2, rg);
    /* levels 5-7 across the eight lanes of the column: the lane with the lower g keeps the sum, the other the difference */
    float p = __shfl_xor_sync(0xffffffffu, s4, 4);
    const float s5 = div_c<FAST>(__fadd_rn(s4, p), s2, r2, rg), d5 = div_c<FAST>(__fsub_rn(p, s4), s2, r2, rg);
    p = __shfl_xor_sync(0xffffffffu, s5, 8);
    const float s6 = div_c<FAST>(__fadd_rn(s5, p), s2, r2, rg), d6 = div_c<FAST>(__fsub_rn(p, s5), s2, r2, rg);
    p = __shfl_xor_sync(0xffffffffu, s6, 16);
    const float s7 = div_c<FAST>(__fadd_rn(s6, p), s2, r2, rg), d7 = div_c<FAST>(__fsub_rn(p, s6), s2, r2, rg);
#pragma unroll
    for (int i = 0; i < 8; i++) coef[i] = d1[i];
#pragma unroll
    for (int i = 0; i < 4; i++) coef[8 + i] = d2[i];
    coef[12] = d3[0]; coef[13] = d3[1]; coef[14] = d4;
    coef[15] = (g & 1) ? d5 : (g & 2) ? d6 : (g & 4) ? d7 : s7;
}

/* One spectral image (128 x 32) by the 256 threads of a CTA: Haar rows + columns (Frame.m:113-153), ordered top-T and packing
 * (Frame.m:165-191); writes the 2 W packed words of the subfingerprint to `words`.  The Haar transform is warp-local — every level
 * runs in registers or through warp shuffles — so the image needs ONE block barrier (between the row and the column pass); the
 * coefficients never go back to shared memory: each thread keeps its 16 and the selection works on registers.  imgT (32 x HS32_LDT
 * floats, 16-byte aligned) and sm are the caller's shared memory; the call ends with a block barrier after the last use of both. */
template <int SRC_LD>
__device__ __forceinline__ void haar_select32_image(const float* __restrict__ img, float* __restrict__ imgT, Select32Smem& sm,
                                                    float* __restrict__ haar_out, uint32_t* __restrict__ words, const int T, const int W) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* bucket = reinterpret_cast<uint32_t*>(imgT);     /* imgT doubles as the bucket of undecided keys once the columns are in registers */
    const int T4 = (T + 3) & ~3;
    float coef[16];
    uint32_t key[16];
    DivRange rg;
    haar32_rows<true, SRC_LD>(img, imgT, rg);
    __syncthreads();
    haar32_cols<true>(imgT, coef, rg);
    auto histogram = [&]() {                             /* per-warp histograms of the exponents of my 16 keys */
#pragma unroll
        for (int e = 0; e < 16; e++) key[e] = __float_as_uint(coef[e]) & 0x7fffffffu;
#pragma unroll
        for (int j = 0; j < 8; j++) sm.hist[wid][lane + 32 * j] = 0;
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 16; e++) atomicAdd(&sm.hist[wid][key[e] >> 23], 1u);
    };
    histogram();
    if (tid < 16) sm.words[tid] = 0;
    if (tid == 0) { sm.nsurv = 0; sm.nbucket = 0; sm.rank_sum = 0; }
#pragma unroll
    for (int j = 0; j < 3; j++) sm.digit_hist[j][tid] = 0;
    for (int i = T + tid; i < T4; i += HS32_THREADS) sm.surv_key[i] = 0;     /* padding of the 128-bit loads of the ranking */
    if (__syncthreads_or(rg.bad())) {                    /* a dividend outside the range of the short division: redo the image with checked divisions */
        haar32_rows<false, SRC_LD>(img, imgT, rg);
        __syncthreads();
        haar32_cols<false>(imgT, coef, rg);
        histogram();
        __syncthreads();
    }
    /* from here on every warp has read its columns: imgT may become the bucket */
    const int cl = lane & 3, g = lane >> 2, col = 4 * wid + cl;
    /* row position of each of my coefficients in the ordered output (flat index = 32 * position + column) */
    const uint32_t pos_last = (g & 1) ? 4 + (g >> 1) : (g & 2) ? 2 + (g >> 2) : (g & 4) ? 1 : 0;
    auto flat_idx = [&](int e) -> uint32_t {
        const uint32_t pos = e < 8 ? 64 + 8 * g + e : e < 12 ? 32 + 4 * g + (e - 8) : e < 14 ? 16 + 2 * g + (e - 12) : e == 14 ? 8 + g : pos_last;
        return pos * 32 + col;
    };
    if (haar_out) {
        float* o = haar_out;
#pragma unroll
        for (int e = 0; e < 16; e++) o[flat_idx(e)] = coef[e];
    }

    /* ---- ordered top-T (Frame.m:165-191): threshold = T-th largest |v| as an integer key ---- */
    {
        /* thread t owns exponent bin t: suffix sums S[t] = #keys with exponent >= t; the threshold's exponent is the largest t with S[t] >= T */
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < HS32_THREADS / 32; w++) tot += sm.hist[w][tid];
        uint32_t suf = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_down_sync(0xffffffffu, suf, d); if (lane + d < 32) suf += o; }
        if (lane == 0) sm.warp_tot[wid] = suf;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < HS32_THREADS / 32; w++) if (w > wid) suf += sm.warp_tot[w];
        if (suf >= (uint32_t)T && suf - tot < (uint32_t)T) { sm.expo = (uint32_t)tid; sm.above = suf - tot; }
    }
    __syncthreads();
    const uint32_t expo = sm.expo;
    {
        /* compact the keys that share the threshold's exponent: only their mantissas are still undecided */
        uint32_t mine = 0;
#pragma unroll
        for (int e = 0; e < 16; e++) mine += (key[e] >> 23) == expo;
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
        uint32_t base = 0;
        if (lane == 31 && incl) base = atomicAdd(&sm.nbucket, incl);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
#pragma unroll
        for (int e = 0; e < 16; e++) if ((key[e] >> 23) == expo) bucket[base++] = key[e];
    }
    __syncthreads();
    /* the remaining 23 bits of the threshold, 8 + 8 + 7 at a time: all threads histogram the digit of the bucket keys that still
     * match, then EVERY warp finds the digit's bin for itself — three short phases instead of a 23-step bisection by one warp */
    const uint32_t nb = sm.nbucket;
    uint32_t need = (uint32_t)T - sm.above, thr = expo << 23, n_eq;           /* need: rank of the threshold among the candidates, >= 1 */
    {
        uint32_t d, above;
        for (uint32_t i = tid; i < nb; i += HS32_THREADS) atomicAdd(&sm.digit_hist[0][(bucket[i] >> 15) & 255u], 1u);
        __syncthreads();
        warp_find_bin256(sm.digit_hist[0], need, d, above, n_eq);
        thr |= d << 15; need -= above;
        for (uint32_t i = tid; i < nb; i += HS32_THREADS) { const uint32_t k = bucket[i]; if ((k >> 15) == (thr >> 15)) atomicAdd(&sm.digit_hist[1][(k >> 7) & 255u], 1u); }
        __syncthreads();
        warp_find_bin256(sm.digit_hist[1], need, d, above, n_eq);
        thr |= d << 7; need -= above;
        for (uint32_t i = tid; i < nb; i += HS32_THREADS) { const uint32_t k = bucket[i]; if ((k >> 7) == (thr >> 7)) atomicAdd(&sm.digit_hist[2][k & 127u], 1u); }
        __syncthreads();
        warp_find_bin256(sm.digit_hist[2], need, d, above, n_eq);
        thr |= d; need -= above;
    }
    /* keys above the threshold all survive; of the n_eq equal to it, the first need_eq in flat-index order (Q9) */
    const uint32_t need_eq = need;
    uint32_t cut = 0xffffffffu;
    if (n_eq > need_eq) {                                                      /* block-uniform; rare (exact magnitude ties at the threshold) */
        /* smallest cut with #(ties with flat index < cut) >= need_eq, by bisection on the 13 index bits */
        if (tid < 16) sm.steps[tid] = 0;
        __syncthreads();
        uint32_t m = 0;
        for (int bit = 12; bit >= 0; --bit) {
            const uint32_t cand = m | (1u << bit);
            uint32_t c = 0;
#pragma unroll
            for (int e = 0; e < 16; e++) c += (key[e] == thr && flat_idx(e) < cand) ? 1u : 0u;
            c = __reduce_add_sync(0xffffffffu, c);
            if (lane == 0 && c) atomicAdd(&sm.steps[bit], c);
            __syncthreads();
            if (sm.steps[bit] < need_eq) m = cand;
        }
        cut = m + 1;
    }
    /* survivors take a slot each (any order: their rank is computed below from key and index) */
#pragma unroll
    for (int e = 0; e < 16; e++) {
        if (key[e] >= thr) {
            if (key[e] > thr || flat_idx(e) < cut) {
                const uint32_t slot = atomicAdd(&sm.nsurv, 1u);
                if (slot < 256u) {
                    sm.surv_key[slot] = key[e];
                    sm.surv_idx[slot] = flat_idx(e) | ((coef[e] > 0.0f ? 1u : 0u) << 16) | ((coef[e] < 0.0f ? 1u : 0u) << 17);
                }
            }
        }
    }
    __syncthreads();
    /* rank by counting.  Fast form: rank = #keys greater than mine, four keys per 128-bit load; it is a permutation exactly when
     * no two survivors have equal keys, which the sum of the ranks tells (ties share the lower rank, so the sum falls short). */
    uint32_t my_rank = 0, my_is = 0;
    if (tid < T) {
        const uint32_t ks = sm.surv_key[tid];
        my_is = sm.surv_idx[tid];
        for (int j = 0; j < T4; j += 4) {
            const uint4 k4 = *reinterpret_cast<const uint4*>(&sm.surv_key[j]);
            my_rank += (k4.x > ks) + (k4.y > ks) + (k4.z > ks) + (k4.w > ks);
        }
    }
    {
        const uint32_t rs = __reduce_add_sync(0xffffffffu, my_rank);
        if (lane == 0 && rs) atomicAdd(&sm.rank_sum, rs);
    }
    __syncthreads();
    if (sm.rank_sum != (uint32_t)(T * (T - 1) / 2)) {                         /* equal keys among the survivors: the index breaks the tie (Q9) */
        if (tid < T) {
            const uint32_t ks = sm.surv_key[tid], idx = my_is & 0xffffu;
            my_rank = 0;
            for (int j = 0; j < T; j++) {
                const uint32_t kj = sm.surv_key[j], ij = sm.surv_idx[j] & 0xffffu;
                my_rank += (kj > ks || (kj == ks && ij < idx)) ? 1u : 0u;
            }
        }
    }
    if (tid < T) {
        if (my_is & (1u << 16)) atomicOr(&sm.words[my_rank >> 5], 1u << (my_rank & 31));
        if (my_is & (1u << 17)) atomicOr(&sm.words[W + (my_rank >> 5)], 1u << (my_rank & 31));
    }
    __syncthreads();
    if (tid < 2 * W) words[tid] = sm.words[tid];
    __syncthreads();
}

/* One CTA per spectral image; about 30 KB of shared memory and 64 registers: four CTAs per SM hide each other's barriers. */
__global__ void __launch_bounds__(HS32_THREADS, 4)
haar_select32_kernel(const float* __restrict__ images, float* __restrict__ haar_out, uint32_t* __restrict__ words,
                     const int T, const int W, const uint32_t total_frames) {
    __shared__ __align__(16) float imgT[32 * HS32_LDT];
    __shared__ Select32Smem sm;
    for (uint32_t f = blockIdx.x; f < total_frames; f += gridDim.x)
        haar_select32_image<0>(images + (size_t)f * LBAD_ROWS_PER_FRAME * 32, imgT, sm, haar_out ? haar_out + (size_t)f * LBAD_ROWS_PER_FRAME * 32 : nullptr,
                               words + (size_t)f * 2 * W, T, W);
}

/* -------------------------------------------------------------------------------------------- fused path ---- */

/* sum of v[a..b) with four interleaved partial sums */
__device__ __forceinline__ float seg_sum(const float* __restrict__ v, uint32_t a, const uint32_t b) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    for (; a + 4 <= b; a += 4) { s0 += v[a]; s1 += v[a + 1]; s2 += v[a + 2]; s3 += v[a + 3]; }
    if (a + 2 <= b) { s0 += v[a]; s1 += v[a + 1]; a += 2; }
    if (a < b) s2 += v[a];
    return (s0 + s1) + (s2 + s3);
}

/* same sum with compile-time bounds on the length, LMIN <= len <= L: fully unrolled, the first LMIN terms unconditional, the rest predicated */
template <int LMIN, int L>
__device__ __forceinline__ float seg_sum_static(const float* __restrict__ v, const uint32_t a, const uint32_t len) {
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int i = 0; i < L; i++) {
        if (i < LMIN || (uint32_t)i < len) { if (i & 1) s1 += v[a + i]; else s0 += v[a + i]; }
    }
    return s0 + s1;
}
constexpr int STATIC_HALF0 = 9, STATIC_HALF1 = 25;     /* longest half-band of bands 0..15 / 16..31 in the reference-default table */
#ifndef LBAD_BANDSUM_MIN
#define LBAD_BANDSUM_MIN 0          /* measured: the unconditional head saves 20 instructions per window but narrows the choice of cuts (57 instead of 50 wavefronts) */
#endif
constexpr int STATIC_MIN0 = LBAD_BANDSUM_MIN ? 3 : 0, STATIC_MIN1 = LBAD_BANDSUM_MIN ? 9 : 0;        /* ... and the shortest the cuts are allowed to make them (band widths there: >= 6 / >= 18) */

constexpr int FUSED_WARPS = 8;
constexpr int FUSED_THREADS = FUSED_WARPS * 32;
constexpr int SCR_LDF = 36;                  /* transpose scratch row stride (floats): LDS.128-aligned and conflict-free */
constexpr int SCR_LD2 = 34;                  /* pair transposition (AOS): row stride in (re, im) pairs — 272 bytes: STS.64 rows and LDS.128 columns both conflict-free */

struct FusedSmemLayout {
    uint32_t samples_bytes, total_bytes;
    uint32_t off_scratch, off_tw1, off_tw2, off_bar;
};
/* groups: warp groups of FUSED_WARPS warps per CTA, each with its own sample buffer, barrier and scratch; carried: the carried-transform
 * kernel reads 17 rows of the pass-1 table and 16 of the real-split table (8.25 KB instead of 16 KB) */
static FusedSmemLayout fused_layout(uint32_t span_floats, bool aos = false, uint32_t groups = 1, bool carried = false) {
    FusedSmemLayout L;
    L.samples_bytes = (span_floats * 4 + 127) & ~127u;
    uint32_t o = L.samples_bytes * groups;
    L.off_scratch = o; o += groups * FUSED_WARPS * 32 * (aos ? SCR_LD2 * 2 : SCR_LDF) * 4;      /* one float component at a time: 4.5 KB per warp; whole pairs (AOS): 8.5 KB */
    L.off_tw1 = o;     o += carried ? 17 * 32 * 8 : 32 * 32 * 8;
    L.off_tw2 = o;     o += carried ? 16 * 32 * 8 : 32 * 32 * 8;
    L.off_bar = o;     o += 16 * groups;
    L.total_bytes = o;
    return L;
}

/* R: lanes per window in pass 1 = points of the pass-2 DFT; window = 64 R samples, 32/R windows per warp (see lbad_math.cuh).
 * STATIC_RANGE (R = 32 only): the band table is the reference-default one (bins [86, 759)), so the rows of 32 bins that need
 * the real split are k2 = 2..23 at compile time; otherwise the range is a (warp-uniform) run-time value. */
/* SUBS: CTA iterations per frame.  1 for throughput (a frame per iteration: every warp carries its half transform over 16 windows).
 * 8 for latency, when a call brings fewer frames than half the CTAs the device holds (a single clip, LBAudioDetectiveProcessPCM):
 * a frame is then spread over eight CTAs of 16 windows, two per warp — same windows, same arithmetic (a carried half transform and a
 * recomputed one are the same instructions on the same samples), one eighth of the serial chain. */
/* AOS: the transposition between the two passes moves whole (re, im) pairs — 32 STS.64 + 16 LDS.128, ONE round trip through an 8.5 KB
 * per-warp scratch, and the loaded registers are (re, im) pairs already, so pass 2 runs packed from its first stage — instead of one
 * component at a time (64 STS.32 + 16 LDS.128 in two round trips through 4.5 KB).  Same values, same arithmetic per element. */
/* GROUPS = 2: ONE CTA per SM of two independent 8-warp groups, each working on its own frame (own sample buffer, mbarrier and named
 * barrier) and sharing only the twiddle tables: what two CTAs per SM did, in the shared memory that two CTAs with pair scratch no longer fit. */
template <int R, bool STATIC_RANGE, bool CARRY, int SUBS = 1, bool AOS = false, int GROUPS = 1>
#ifndef LBAD_G1_MINBLOCKS
#define LBAD_G1_MINBLOCKS 2          /* A/B knob: 1 lets the one-group variants use up to 255 registers */
#endif
__global__ void __launch_bounds__(FUSED_THREADS * GROUPS, GROUPS == 1 ? LBAD_G1_MINBLOCKS : 1)
bands_fused_kernel(const float* __restrict__ pcm, float* __restrict__ images, const float4* __restrict__ g_tw1, const float4* __restrict__ g_tw2,
                   const Geo g, const BandTable bt, const FusedSmemLayout L, const uint32_t span_floats,
                   const uint32_t total_frames, const int use_tma, const uint32_t frame0) {
    static_assert(!STATIC_RANGE || R == 32, "the compile-time band rows belong to the 2048-sample window");
    static_assert(!CARRY || R == 32, "half transforms are carried between windows only in the one-window-per-warp layout");
    static_assert(SUBS == 1 || CARRY, "frames are split between CTAs only in the carried layout");
    constexpr uint32_t UNIT_ROWS = LBAD_ROWS_PER_FRAME / SUBS;      /* windows per CTA iteration */
    constexpr int S = 32 / R;                                /* windows per warp */
    constexpr int M = 32 * R;                                /* complex points per window = bins of the half spectrum */
    static_assert(GROUPS == 1 || (CARRY && SUBS == 1), "two groups per CTA only in the throughput form of the carried kernel");
    extern __shared__ __align__(128) unsigned char smem[];
    const int grp = GROUPS == 1 ? 0 : (int)(threadIdx.x / FUSED_THREADS);               /* warp group of this thread */
    const int tid = GROUPS == 1 ? (int)threadIdx.x : (int)(threadIdx.x % FUSED_THREADS), lane = tid & 31, wid = tid >> 5;   /* thread / warp index inside the group */
    float*  samples = reinterpret_cast<float*>(smem + (size_t)grp * L.samples_bytes);
    float4* tw1 = reinterpret_cast<float4*>(smem + L.off_tw1);     /* [p/2][lane]: twiddles of register positions p, p+1 */
    float4* tw2 = reinterpret_cast<float4*>(smem + L.off_tw2);     /* R < 32: [k2/2][lane]: (cos, sin) of bins lane+32k2, lane+32(k2+1); R == 32: float2 [k2][lane] */
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.off_bar) + 2 * grp;
    float* scr = reinterpret_cast<float*>(smem + L.off_scratch) + (grp * FUSED_WARPS + wid) * (32 * (AOS ? SCR_LD2 * 2 : SCR_LDF));
    float* vbuf = scr;                                       /* band scratch (S x M = 1024 floats) aliases the transpose scratch */
    auto group_sync = [&]() {                                /* all threads of this warp group */
        if constexpr (GROUPS == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(FUSED_THREADS) : "memory");
    };

    {
        constexpr int TW1_ENTRIES = CARRY ? 17 * 16 : 512, TW2_ENTRIES = CARRY ? 16 * 16 : 512;      /* float4 entries actually read (see fused_layout) */
        for (int i = threadIdx.x; i < TW1_ENTRIES; i += FUSED_THREADS * GROUPS) tw1[i] = g_tw1[i];
        for (int i = threadIdx.x; i < TW2_ENTRIES; i += FUSED_THREADS * GROUPS) tw2[i] = g_tw2[i];
    }
    if (tid == 0 && use_tma) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();

    /* round 0 covers bands 0..15, round 1 bands 16..31; lanes 2b and 2b+1 split band b (resp. 16+b) in halves */
    uint32_t ra0, rb0, ra1, rb1;
    {
        const int b0 = lane >> 1, b1 = 16 + (lane >> 1);
        const uint32_t l0 = bt.klow[b0], h0 = bt.khigh[b0], m0 = bt.split[b0];
        const uint32_t l1 = bt.klow[b1], h1 = bt.khigh[b1], m1 = bt.split[b1];
        ra0 = (lane & 1) ? m0 : l0; rb0 = (lane & 1) ? h0 : m0;
        ra1 = (lane & 1) ? m1 : l1; rb1 = (lane & 1) ? h1 : m1;
    }
    const int my_band = (lane >> 1) + ((lane & 1) ? 16 : 0);
    const float divisor = bt.divisor[my_band];
    const int k2lo = STATIC_RANGE ? 2 : (int)(g.kmin >> 5), k2hi = STATIC_RANGE ? 23 : (int)((g.kmax - 1) >> 5);
    const uint32_t hop = g.stride;
    const uint32_t bytes = span_floats * 4;
    const float scale_m1 = g.inv_pos_scale - 1.0f;
    const int my_win = lane / R, n2 = lane % R;              /* pass 1: which of the warp's windows this lane loads, and its column */

    auto frame_src = [&](uint32_t ul) -> const float* {                        /* ul: unit (frame x SUBS + part) index inside this launch's slab */
        const uint32_t f = frame0 + ul / SUBS, part = ul % SUBS;
        const uint32_t clip = f / g.frames_per_clip, fr = f % g.frames_per_clip;
        return pcm + (uint64_t)clip * g.clip_stride + ((uint64_t)fr * LBAD_ROWS_PER_FRAME + part * UNIT_ROWS) * hop;   /* m:262-290 */
    };

    uint32_t f = blockIdx.x * GROUPS + grp, parity = 0;
    if (f < total_frames && use_tma && tid == 0) { mbar_arrive_expect_tx(bar, bytes); bulk_copy_g2s(samples, frame_src(f), bytes, bar); }

    for (; f < total_frames; f += gridDim.x * GROUPS) {
        if (use_tma) { mbar_wait(bar, parity); parity ^= 1; }
        else {
            const float* src = frame_src(f);
            for (uint32_t i = tid; i < span_floats; i += FUSED_THREADS) samples[i] = src[i];
            group_sync();
        }

        /* ---- the frame's 128 windows, S at a time per warp, round-robin over the warps: FFT -> bands -> image rows ---- */
        constexpr int ITERS = (int)UNIT_ROWS / (FUSED_WARPS * S);
        /* CARRY (R == 32 and hop == 64 samples): a warp takes ITERS CONSECUTIVE windows, one hop = one n1 step apart, so the
         * 16-point transform of a window's odd-n1 inputs is carried over as the even-n1 transform of the next window (dft16 /
         * dit32_combine in lbad_math.cuh): half the sample loads and 48 instead of 80 butterflies in pass 1. */
        float2 carry[16];
        const float2* tw1h = reinterpret_cast<const float2*>(tw1);               /* CARRY: [q][lane] = exp(-2 pi i n2 k1 / M), k1 = bitrev4(q); row 16 = exp(-2 pi i 16 n2 / M) */
        /* The pass-1 twiddle exp(-2 pi i n2 k1 / M) does not depend on the window either, so what is carried is P = O x twiddle:
         * z[k1] x tw[k1] = P_prev[k1] + W^k1 P[k1], and z[k1+16] x tw[k1+16] = (P_prev[k1] - W^k1 P[k1]) x exp(-2 pi i 16 n2 / M) —
         * 16 table entries per window instead of 32 (the kernel is bound by shared-memory wavefronts), the same number of products. */
        auto half_transform = [&](const float* w, float2 (&h)[16]) {             /* inputs z[32 n1 + lane], n1 = 1, 3, .., 31 of the window at w */
#pragma unroll
            for (int m = 0; m < 16; m++) h[m] = *reinterpret_cast<const float2*>(w + 2 * (32 * (2 * m + 1) + lane));
            dft16(h);
            /* one table entry per output (17 LDS.64 per window with the entry of the odd outputs).  Deriving eleven of the sixteen powers
             * of the lane's root from four loaded ones — 24 wavefronts less, 22 packed multiplies more — was measured: 34.0 against 33.5 ms;
             * the FMA pipe is as busy as the shared-memory pipe by now, and the twiddles lose half a digit. */
#pragma unroll
            for (int q = 1; q < 16; q++) h[q] = cmul(h[q], tw1h[q * 32 + lane]);       /* q = 0 is k1 = 0: twiddle 1 */
        };
        if constexpr (CARRY) half_transform(samples + (size_t)(wid * ITERS) * hop - hop, carry);     /* even n1 of the first window = odd n1 of the one before */
#ifndef LBAD_WINDOW_UNROLL
#define LBAD_WINDOW_UNROLL 2
#endif
        /* the carried kernel handles two windows per trip: the half transform carried out of the first is the one carried into the second,
         * so the 32 register moves that would hand it over disappear (35.4 against 36.3 ms); four windows per trip outgrow the instruction
         * cache (40.1 ms).  A/B knob: build.py --variant NAME -DLBAD_WINDOW_UNROLL=n */
        constexpr int WINDOW_UNROLL = CARRY ? LBAD_WINDOW_UNROLL : 1;
#pragma unroll WINDOW_UNROLL
        for (int it = 0; it < ITERS; it++) {
            const int row0 = CARRY ? wid * ITERS + it : (wid + it * FUSED_WARPS) * S;
            const float* win = samples + (size_t)(row0 + my_win) * hop;
            float2 z[32];
            if constexpr (CARRY) {
                float2 odd[16];
                half_transform(win, odd);
                dit32_combine(carry, odd, z);                                   /* over n1; position p holds k1 = bitrev5(p), already twiddled for even p */
                const float2 om = tw1h[16 * 32 + lane];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    z[2 * q + 1] = cmul(z[2 * q + 1], om);
                    carry[q] = odd[q];
                }
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++)                                 /* vDSP_ctoz (m:353): z[n] = x[2n] + i x[2n+1], n = R n1 + n2 */
                    z[n1] = *reinterpret_cast<const float2*>(win + 2 * (R * n1 + n2));
                fft32(z);                                                       /* over n1; position p holds k1 = bitrev5(p) */
#pragma unroll
                for (int p = 0; p < 32; p += 2) {                               /* x exp(-2 pi i n2 k1 / M) */
                    const float4 w = tw1[(p >> 1) * 32 + lane];
                    z[p] = cmul(z[p], make_float2(w.x, w.y));
                    z[p + 1] = cmul(z[p + 1], make_float2(w.z, w.w));
                }
            }
            if constexpr (AOS) {
                /* 32x32 transpose of (re, im) pairs: row k1 = bitrev5(p) receives this lane's pair at column lane; lane k1 then reads its row */
                float2* scr2 = reinterpret_cast<float2*>(scr);
#pragma unroll
                for (int p = 0; p < 32; p++) scr2[bitrev5(p) * SCR_LD2 + lane] = z[p];
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const float4 t = *reinterpret_cast<const float4*>(&scr2[lane * SCR_LD2 + 2 * q]);
                    z[2 * q] = make_float2(t.x, t.y); z[2 * q + 1] = make_float2(t.z, t.w);
                }
                fft32_tail<R>(z);                                                   /* over n2, per window; position s R + q holds Z_s[lane + 32 bitrevR(q)] */
            } else {
            /* 32x32 transpose through shared memory, one component at a time: lane k1 ends up with A_s[n2][k1] at position s R + n2 */
            float zx[32], zy[32];
#pragma unroll
            for (int p = 0; p < 32; p++) scr[bitrev5(p) * SCR_LDF + lane] = z[p].x;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 t = *reinterpret_cast<const float4*>(&scr[lane * SCR_LDF + 4 * q]);
                zx[4 * q] = t.x; zx[4 * q + 1] = t.y; zx[4 * q + 2] = t.z; zx[4 * q + 3] = t.w;
            }
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 32; p++) scr[bitrev5(p) * SCR_LDF + lane] = z[p].y;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 t = *reinterpret_cast<const float4*>(&scr[lane * SCR_LDF + 4 * q]);
                zy[4 * q] = t.x; zy[4 * q + 1] = t.y; zy[4 * q + 2] = t.z; zy[4 * q + 3] = t.w;
            }
            fft32_tail_soa<R>(zx, zy, z);                                       /* over n2, per window; position s R + q holds Z_s[lane + 32 bitrevR(q)] */
            }
            __syncwarp();                                                       /* scratch is about to be reused as vbuf */
            const int src_lane = (32 - lane) & 31;
            if constexpr (R == 32) {
                /* rows k2 and 31 - k2 are mirror images (bin k of lane l <-> bin 1024 - k of lane 32 - l): one evaluation of the
                 * real split yields both, and the lane that did it stores both energies.  Lane 0 is its own mirror one row up
                 * (k = 32 k2 <-> 32 (32 - k2)), so its row 16 (bin 512, self-mirrored) is done separately below. */
                auto row_needed = [&](int r) -> bool { return STATIC_RANGE ? (r >= 2 && r <= 23) : (r >= k2lo && r <= k2hi); };
#ifndef LBAD_SPLIT_TABLE
#define LBAD_SPLIT_TABLE 0
#endif
                /* the twiddle exp(i 2 pi k / N) of bin k = lane + 32 k2 is the lane's factor (row 0 of the table: ONE load per window) times
                 * the row's compile-time factor (lbad_math.cuh real_split_pair_row): the kernel is bound by the shared-memory pipe, a table
                 * entry per row cost 28 wavefronts per window, and with packed arithmetic the rotation costs no instruction more than the load
                 * it replaces (LBAD_SPLIT_TABLE=1 keeps the table form for A/B) */
                float2 wl = make_float2(1.0f, 0.0f);
                if constexpr (!LBAD_SPLIT_TABLE) wl = reinterpret_cast<const float2*>(tw2)[lane];
#pragma unroll
                for (int k2 = 0; k2 < 16; k2++) {
                    const bool need_lo = row_needed(k2), need_hi = row_needed(31 - k2) || (k2 > 0 && row_needed(32 - k2));   /* warp-uniform */
                    if (need_lo || need_hi) {
                        const int p = bitrev5(k2), pp = bitrev5(31 - k2), p0 = bitrev5((32 - k2) % 32);
                        /* Z[1024 - k] lives in lane 32 - lane, row 31 - k2 ... except for lane 0: own register, row 32 - k2.  Lane 0 is
                         * read by nobody but itself, so it offers that register and the exception costs nothing after the shuffle. */
                        const float2 offer = lane == 0 ? z[p0] : z[pp];
                        float2 pz;
                        pz.x = __shfl_sync(0xffffffffu, offer.x, src_lane);
                        pz.y = __shfl_sync(0xffffffffu, offer.y, src_lane);
                        const int k = k2 * 32 + lane;
                        float2 lo, hi;
                        if constexpr (LBAD_SPLIT_TABLE) {
                            const float2 w = reinterpret_cast<const float2*>(tw2)[k2 * 32 + lane];   /* (cos, sin) of 2 pi k / N: one conflict-free LDS.64 */
                            real_split_pair_2x(z[p], pz, w.x, w.y, lo, hi);
                        } else real_split_pair_rows(k2, z[p], pz, wl, lo, hi);
                        if (k2 == 0 && lane == 0) { lo.x = 2.0f * (z[p].x + z[p].y); lo.y = 2.0f * (z[p].x - z[p].y); }       /* DC / packed Nyquist */
                        if (need_lo) vbuf[k] = bin_energy_raw(lo.x, lo.y, scale_m1);                                         /* finiteness is checked on the band sum */
                        if (need_hi && (k2 > 0 || lane > 0)) vbuf[1024 - k] = bin_energy_raw_conj(hi.x, hi.y, scale_m1);      /* (a row without a mirror: hi is dead code) */
                    }
                }
                if (row_needed(16) && lane == 0) vbuf[512] = bin_energy_raw(2.0f * z[bitrev5(16)].x, -2.0f * z[bitrev5(16)].y, scale_m1);   /* w = -i, partner = itself */
            } else {
    #pragma unroll
                for (int k2 = 0; k2 < R; k2 += 2) {
                    if ((STATIC_RANGE && k2 >= 2 && k2 <= 23) || (!STATIC_RANGE && k2 + 1 >= k2lo && k2 <= k2hi)) {   /* warp-uniform */
                        const float4 w = tw2[(k2 >> 1) * 32 + lane];                /* (cos, sin) of 2 pi k / N for k2 and k2+1 */
    #pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int kk = k2 + h;
    #pragma unroll
                            for (int s = 0; s < S; s++) {
                                const int p = s * R + bitrevR<R>(kk), pp = s * R + bitrevR<R>(R - 1 - kk), p0 = s * R + bitrevR<R>((R - kk) % R);
                                const float2 offer = lane == 0 ? z[p0] : z[pp];     /* Z[M - k] lives in lane 32-lane, k2' = R-1-k2 ... except lane 0: own register k2' = R-k2 */
                                float2 pz;
                                pz.x = __shfl_sync(0xffffffffu, offer.x, src_lane);
                                pz.y = __shfl_sync(0xffffffffu, offer.y, src_lane);
                                float xr, xi;
                                real_split_2x(z[p], pz, h ? w.z : w.x, h ? w.w : w.y, xr, xi);
                                if (kk == 0 && lane == 0) { xr = 2.0f * (z[p].x + z[p].y); xi = 2.0f * (z[p].x - z[p].y); }   /* DC / packed Nyquist */
                                vbuf[s * M + kk * 32 + lane] = bin_energy_raw(xr, xi, scale_m1);      /* finiteness is checked on the band sum */
                            }
                        }
                    }
                }
            }
            __syncwarp();
            /* band sums, m:379-405: two lanes per band (each sums half of the band's bins), 16 bands per round */
#pragma unroll
            for (int s = 0; s < S; s++) {
                const float* v = vbuf + s * M;
                float sa, sb;
                if (STATIC_RANGE) { sa = seg_sum_static<STATIC_MIN0, STATIC_HALF0>(v, ra0, rb0 - ra0); sb = seg_sum_static<STATIC_MIN1, STATIC_HALF1>(v, ra1, rb1 - ra1); }
                else              { sa = seg_sum(v, ra0, rb0); sb = seg_sum(v, ra1, rb1); }
                sa += __shfl_xor_sync(0xffffffffu, sa, 1);
                sb += __shfl_xor_sync(0xffffffffu, sb, 1);
                float tot = (lane & 1) ? sb : sa;                              /* even lane: band lane/2, odd lane: band 16 + lane/2 */
                if (!(tot <= 3.402823466e+38f)) {                               /* a non-finite term (m:398-401 skips it): re-sum this band filtered */
                    tot = 0.0f;
                    for (uint32_t k = bt.klow[my_band]; k < bt.khigh[my_band]; k++) { const float e = v[k]; if (e <= 3.402823466e+38f) tot += e; }
                }
                /* the 32 lanes fill one 128-byte line of the image row */
                images[((size_t)f * UNIT_ROWS + row0 + s) * 32 + my_band] = __fdiv_rn(tot, divisor);      /* f counts units: unit f holds rows [f UNIT_ROWS, (f + 1) UNIT_ROWS) of the slab */
            }
            __syncwarp();
        }
        group_sync();                                                           /* every warp of the group is done with the samples */
        const uint32_t fnext = f + gridDim.x * GROUPS;
        if (use_tma && tid == 0 && fnext < total_frames) { mbar_arrive_expect_tx(bar, bytes); bulk_copy_g2s(samples, frame_src(fnext), bytes, bar); }
    }
}

}  // namespace lbad

/* ================================================================================================== host ==== */

using namespace lbad;

struct lbadcu_plan {
    lbadcu_geometry geo;
    Geo g;
    BandTable bt;
    int device = 0;
    cudaStream_t stream = nullptr, copy_streams[3] = {nullptr, nullptr, nullptr};
    float2 *d_tw_m = nullptr, *d_tw_n = nullptr; float4 *d_tw1 = nullptr, *d_tw2 = nullptr;
    bool static_range = false;
    float* d_scratch[4] = {nullptr, nullptr, nullptr, nullptr}; size_t scratch_frames[4] = {0, 0, 0, 0};   /* spectral images: slot 0 for device-API calls, 1..3 for the host pipeline's chunk buffers */
    /* a slot's images live between the two kernels of one call: a call on another stream than the slot's previous user waits for that
     * one's second kernel (calls on one detective are serialised on the device, whatever streams they are enqueued on) */
    cudaEvent_t scratch_done[4] = {nullptr, nullptr, nullptr, nullptr}; cudaStream_t scratch_stream[4] = {nullptr, nullptr, nullptr, nullptr}; bool scratch_used[4] = {false, false, false, false};
    int16_t* d_chunk_i16[3] = {nullptr, nullptr, nullptr}; size_t chunk_i16 = 0;
    float* d_chunk_pcm[3] = {nullptr, nullptr, nullptr}; uint32_t* d_chunk_words[3] = {nullptr, nullptr, nullptr};
    size_t chunk_pcm_floats = 0, chunk_words = 0;
    bool fused_ok = false; int sm_count = 0; size_t smem_optin = 0;
    uint64_t launches = 0;
    LaunchTimer timer, timer2;      /* FFT+bands kernel / Haar+select kernel */
    int stage_mode = -1;      /* -1 auto, 0 plain loads, 1 TMA (env LBAD_STAGE=ldg|tma) */
    bool transform_generic = false;      /* env LBAD_TRANSFORM=generic: lbadcu_transform_images_host uses the any-geometry Haar/select kernel */
    uint32_t slab_frames_cap = 1u << 18;
    uint32_t chunk_clips_override = 0;   /* env LBAD_CHUNK_CLIPS: clips per upload chunk of the host pipeline (tests: many chunks from few clips); 0 = by size */
    struct { const void* fn; uint32_t smem; int per_sm; } occ_cache[8] = {};      /* per kernel variant: opt-in shared memory set, resident CTAs per SM */
    int force_subs = 0;                  /* env LBAD_SUBFRAMES=1|2|8: pin the CTA iterations per frame (tests); 0 = by frame count */
    int pair_transpose = -1;             /* env LBAD_TRANSPOSE=pairs|components: the AOS / component-wise transposition of the carried kernel; -1 = default (pairs) */
    int force_groups = 0;                /* env LBAD_GROUPS=1|2: warp groups per CTA of the pair-transposition kernel; 0 = default (2) */
    float *d_score = nullptr, *h_score = nullptr;      /* lbadcu_compare_pcm_host: the match on the device and its pinned landing place */
};

extern "C" const char* lbadcu_last_error(void) { return g_err; }
/* the message of another thread's failure, handed to the thread that reports it (the buffer is per thread) */
extern "C" void lbadcu_set_last_error(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg ? msg : ""); }

extern "C" int lbadcu_device_available(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); return LBAD_ERR_NODEVICE; }
    return LBAD_OK;
}

extern "C" int lbadcu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

/* Makes `device` current for the calling thread and reports the one that was (for lbadcu_pop_device); device < 0: nothing changes. */
extern "C" int lbadcu_push_device(int device, int* prev) {
    *prev = -1;
    if (device < 0) return LBAD_OK;
    if (lbadcu_device_available() != LBAD_OK) { set_error("no CUDA device available (this library has no CPU fallback)"); return LBAD_ERR_NODEVICE; }
    if (device >= lbadcu_device_count()) { set_error("CUDA device %d does not exist (%d device(s))", device, lbadcu_device_count()); return LBAD_ERR_ARG; }
    int cur = -1;
    LBAD_CUDA_TRY(cudaGetDevice(&cur));
    if (cur != device) { LBAD_CUDA_TRY(cudaSetDevice(device)); *prev = cur; }
    return LBAD_OK;
}
extern "C" void lbadcu_pop_device(int prev) { if (prev >= 0) cudaSetDevice(prev); }
extern "C" int lbadcu_plan_device(const lbadcu_plan* p);

static uint32_t ilog2(uint32_t v) { uint32_t l = 0; while ((1u << l) < v) l++; return l; }

/* Where the two lanes of a band split it.  In round r (bands 16 r .. 16 r + 15) step i of the band sums has lane (b, h) read the
 * energy at start(b, h) + i; the shared-memory wavefronts of that instruction are the largest number of lanes whose addresses share
 * a bank.  With every band cut in the middle the 34 loads of the reference-default table take 77 wavefronts; moving the cuts (both
 * parts still at most lmax long) so that fewer starts coincide mod 32 brings them to about 50.  Deterministic annealing, some
 * milliseconds, once per plan. */
static uint32_t band_sum_wavefronts(const uint32_t* lo, const uint32_t* hi, const uint32_t* split, int r, uint32_t lmax) {
    uint32_t total = 0;
    for (uint32_t i = 0; i < lmax; i++) {
        uint32_t cnt[32] = {0}, worst = 0;
        for (int b = 16 * r; b < 16 * r + 16; b++) {
            if (i < split[b] - lo[b]) { const uint32_t c = ++cnt[(lo[b] + i) & 31]; worst = c > worst ? c : worst; }
            if (i < hi[b] - split[b]) { const uint32_t c = ++cnt[(split[b] + i) & 31]; worst = c > worst ? c : worst; }
        }
        total += worst;
    }
    return total;
}
static void optimise_band_splits(const uint32_t* lo, const uint32_t* hi, uint32_t* split, const uint32_t lmax0, const uint32_t lmax1, const uint32_t lmin0, const uint32_t lmin1) {
    for (int r = 0; r < 2; r++) {
        const uint32_t lmax = r ? lmax1 : lmax0, lmin = r ? lmin1 : lmin0;       /* both parts of every band between lmin and lmax long */
        uint32_t cur[32], best[32];
        for (int b = 0; b < 32; b++) cur[b] = best[b] = split[b];
        uint32_t cw = band_sum_wavefronts(lo, hi, cur, r, lmax), bw = cw;
        uint64_t rng = 0x9E3779B97F4A7C15ull + (uint64_t)r;
        double T = 2.0;
        for (int it = 0; it < 40000; it++) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            const int b = 16 * r + (int)((rng >> 33) % 16);
            uint32_t smin = hi[b] > lo[b] + lmax ? hi[b] - lmax : lo[b], smax = lo[b] + lmax < hi[b] ? lo[b] + lmax : hi[b];
            if (smin < lo[b] + lmin) smin = lo[b] + lmin;
            if (smax > hi[b] - lmin) smax = hi[b] - lmin;
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            const uint32_t s = smin + (uint32_t)((rng >> 33) % (smax - smin + 1)), old = cur[b];
            cur[b] = s;
            const uint32_t w = band_sum_wavefronts(lo, hi, cur, r, lmax);
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            const double u = (double)(rng >> 11) * (1.0 / 9007199254740992.0);
            if (w <= cw || u < exp(((double)cw - (double)w) / T)) cw = w; else cur[b] = old;
            if (cw < bw) { bw = cw; for (int k = 16 * r; k < 16 * r + 16; k++) best[k] = cur[k]; }
            T = T * 0.9999 > 0.05 ? T * 0.9999 : 0.05;
        }
        for (int b = 16 * r; b < 16 * r + 16; b++) split[b] = best[b];
    }
}

extern "C" int lbadcu_plan_create(const lbadcu_geometry* geo, lbadcu_plan** out) {
    *out = nullptr;
    if (lbadcu_device_available() != LBAD_OK) { set_error("no CUDA device available (this library has no CPU fallback)"); return LBAD_ERR_NODEVICE; }
    const uint32_t N = geo->window, B = geo->bands, M = N / 2;
    if (N < LBAD_MIN_WINDOW || N > LBAD_MAX_WINDOW || (N & (N - 1)) || B < 4 || B > LBAD_MAX_BANDS || (B & (B - 1)) || geo->stride == 0 ||
        geo->sublen < 2 || geo->sublen > LBAD_MAX_SUBLEN || (geo->sublen & 1) || geo->sublen / 2 > LBAD_ROWS_PER_FRAME * B) return LBAD_ERR_ARG;
    uint32_t kmin = 0xffffffffu, kmax = 0;
    for (uint32_t b = 0; b < B; b++) {
        if (geo->klow[b] > geo->khigh[b] || geo->khigh[b] > M) return LBAD_ERR_ARG;      /* Q15: band table must stay inside the spectrum */
        if (geo->klow[b] < geo->khigh[b]) { kmin = geo->klow[b] < kmin ? geo->klow[b] : kmin; kmax = geo->khigh[b] > kmax ? geo->khigh[b] : kmax; }
    }
    if (kmin >= kmax) { kmin = 0; kmax = 1; }
    lbadcu_plan* p = new lbadcu_plan();
    Guard<lbadcu_plan> guard(p, lbadcu_plan_destroy);                            /* the CUDA calls below return on failure */
    p->geo = *geo;
    LBAD_CUDA_TRY(cudaGetDevice(&p->device));
    cudaDeviceProp prop; LBAD_CUDA_TRY(cudaGetDeviceProperties(&prop, p->device));
    p->sm_count = prop.multiProcessorCount; p->smem_optin = prop.sharedMemPerBlockOptin;
    Geo& g = p->g;
    g.window = N; g.log2m = ilog2(M); g.stride = geo->stride; g.bands = B; g.pairs = geo->sublen / 2;
    g.words_per_plane = g.pairs <= 64 ? 2 : g.pairs <= 128 ? 4 : 8;
    g.kmin = kmin; g.kmax = kmax; g.inv_pos_scale = 1.0f / geo->pos_scale; g.frames_per_clip = 0; g.clip_stride = 0;
    for (uint32_t b = 0; b < LBAD_MAX_BANDS; b++) { p->bt.klow[b] = b < B ? geo->klow[b] : 0; p->bt.khigh[b] = b < B ? geo->khigh[b] : 0; p->bt.divisor[b] = b < B ? geo->divisor[b] : 1.0f; }
    LBAD_CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 3; i++) LBAD_CUDA_TRY(cudaStreamCreateWithFlags(&p->copy_streams[i], cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) LBAD_CUDA_TRY(cudaEventCreateWithFlags(&p->scratch_done[i], cudaEventDisableTiming));
    /* twiddle tables, evaluated in double and rounded once */
    std::vector<float2> twm(M / 2), twn(M); std::vector<float4> tw1(512), tw2(512);
    for (uint32_t j = 0; j < M / 2; j++) { const double a = 2.0 * M_PI * j / M; twm[j] = make_float2((float)cos(a), (float)-sin(a)); }
    for (uint32_t k = 0; k < M; k++) { const double a = 2.0 * M_PI * k / N; twn[k] = make_float2((float)cos(a), (float)sin(a)); }
    const int Rr = (int)(N / 64);                                                 /* lanes per window in the register-FFT kernel (N = 64 R) */
    for (int pp = 0; pp < 32; pp += 2) for (int l = 0; l < 32; l++) {            /* pass-1 twiddles exp(-2 pi i n2 k1 / M), n2 = lane % R, in register-position order */
        const double a0 = 2.0 * M_PI * (double)((l % Rr) * bitrev5(pp)) / (double)M, a1 = 2.0 * M_PI * (double)((l % Rr) * bitrev5(pp + 1)) / (double)M;
        tw1[(pp >> 1) * 32 + l] = make_float4((float)cos(a0), (float)-sin(a0), (float)cos(a1), (float)-sin(a1));
    }
    if (N == 2048 && geo->stride == 64) {                                         /* the carried-transform kernel: one (cos, -sin) per k1 < 16 and lane, then the lane's exp(-2 pi i 16 n2 / M) */
        float2* t1 = reinterpret_cast<float2*>(tw1.data());
        for (int q = 0; q < 16; q++) for (int l = 0; l < 32; l++) { const double a = 2.0 * M_PI * (double)(l * bitrevR<16>(q)) / (double)M; t1[q * 32 + l] = make_float2((float)cos(a), (float)-sin(a)); }
        for (int l = 0; l < 32; l++) { const double a = 2.0 * M_PI * (double)(16 * l) / (double)M; t1[16 * 32 + l] = make_float2((float)cos(a), (float)-sin(a)); }
    }
    for (int k2 = 0; k2 < 32; k2 += 2) for (int l = 0; l < 32; l++) {            /* real-split twiddles for bins l+32k2 and l+32(k2+1) (rows k2 < R are used) */
        const double a0 = 2.0 * M_PI * (double)(l + 32 * k2) / (double)N, a1 = 2.0 * M_PI * (double)(l + 32 * (k2 + 1)) / (double)N;
        if (Rr == 32) {                                                           /* one (cos, sin) per row and lane, read as 64-bit words */
            float2* t2 = reinterpret_cast<float2*>(tw2.data());
            t2[k2 * 32 + l] = make_float2((float)cos(a0), (float)sin(a0)); t2[(k2 + 1) * 32 + l] = make_float2((float)cos(a1), (float)sin(a1));
        } else tw2[(k2 >> 1) * 32 + l] = make_float4((float)cos(a0), (float)sin(a0), (float)cos(a1), (float)sin(a1));
    }
    LBAD_CUDA_TRY(cudaMalloc(&p->d_tw_m, sizeof(float2) * (M / 2))); LBAD_CUDA_TRY(cudaMalloc(&p->d_tw_n, sizeof(float2) * M));
    LBAD_CUDA_TRY(cudaMalloc(&p->d_tw1, sizeof(float4) * 512)); LBAD_CUDA_TRY(cudaMalloc(&p->d_tw2, sizeof(float4) * 512));
    LBAD_CUDA_TRY(cudaMemcpy(p->d_tw_m, twm.data(), sizeof(float2) * (M / 2), cudaMemcpyHostToDevice));
    LBAD_CUDA_TRY(cudaMemcpy(p->d_tw_n, twn.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
    LBAD_CUDA_TRY(cudaMemcpy(p->d_tw1, tw1.data(), sizeof(float4) * 512, cudaMemcpyHostToDevice));
    LBAD_CUDA_TRY(cudaMemcpy(p->d_tw2, tw2.data(), sizeof(float4) * 512, cudaMemcpyHostToDevice));
    p->static_range = N == 2048 && (kmin >> 5) == 2 && ((kmax - 1) >> 5) == 23;
    for (uint32_t b = 0; b < LBAD_MAX_BANDS; b++) p->bt.split[b] = p->bt.klow[b] + (p->bt.khigh[b] - p->bt.klow[b] + 1) / 2;
    for (uint32_t b = 0; b < B && b < 32; b++) {
        const uint32_t width = geo->khigh[b] - geo->klow[b], half = (width + 1) / 2;
        if (half > (uint32_t)(b < 16 ? STATIC_HALF0 : STATIC_HALF1) || width < 2u * (uint32_t)(b < 16 ? STATIC_MIN0 : STATIC_MIN1)) p->static_range = false;
    }
    if (p->static_range && B == 32) {                                             /* (the run-time-range variants sum with plain loops: any cut would do, the middle stays) */
        static std::mutex cache_mutex; static std::vector<std::pair<std::vector<uint32_t>, std::vector<uint32_t>>> cache;      /* band table -> cuts */
        std::vector<uint32_t> key(p->bt.klow, p->bt.klow + 32); key.insert(key.end(), p->bt.khigh, p->bt.khigh + 32);
        std::lock_guard<std::mutex> lock(cache_mutex);
        const std::vector<uint32_t>* hit = nullptr;
        for (auto& e : cache) if (e.first == key) hit = &e.second;
        if (!hit) {
            std::vector<uint32_t> sp(p->bt.split, p->bt.split + 32);
            optimise_band_splits(p->bt.klow, p->bt.khigh, sp.data(), STATIC_HALF0, STATIC_HALF1, STATIC_MIN0, STATIC_MIN1);
            cache.emplace_back(key, sp); hit = &cache.back().second;
        }
        for (int b = 0; b < 32; b++) p->bt.split[b] = (*hit)[b];
    }
    /* fused path: window 2048, 32 bands, even hop, frame span fits in shared memory */
    const uint64_t span = 127ull * geo->stride + N;
    p->fused_ok = ((N == 2048 || N == 1024 || N == 512 || N == 256) && B == 32 && (geo->stride % 2 == 0) && span * 4 < (1u << 20) &&
                   fused_layout((uint32_t)span).total_bytes <= p->smem_optin &&
                   (!(N == 2048 && geo->stride == 64) || fused_layout((uint32_t)span, true, 2, true).total_bytes <= p->smem_optin));
    const char* st = getenv("LBAD_STAGE");
    p->stage_mode = st ? (strcmp(st, "tma") == 0 ? 1 : strcmp(st, "ldg") == 0 ? 0 : -1) : -1;
    if (const char* sb = getenv("LBAD_SUBFRAMES")) p->force_subs = atoi(sb) == 8 ? 8 : atoi(sb) == 2 ? 2 : atoi(sb) == 1 ? 1 : 0;
    if (const char* gp = getenv("LBAD_GROUPS")) p->force_groups = atoi(gp) == 1 ? 1 : atoi(gp) == 2 ? 2 : 0;
    if (const char* tp = getenv("LBAD_TRANSPOSE")) p->pair_transpose = strcmp(tp, "pairs") == 0 ? 1 : strcmp(tp, "components") == 0 ? 0 : -1;
    if (const char* tf = getenv("LBAD_TRANSFORM")) p->transform_generic = strcmp(tf, "generic") == 0;
    if (const char* cc = getenv("LBAD_CHUNK_CLIPS")) { const unsigned long v = strtoul(cc, nullptr, 10); if (v >= 1 && v <= (1u << 20)) p->chunk_clips_override = (uint32_t)v; }
    if (const char* sf = getenv("LBAD_SLAB_FRAMES")) { const unsigned long v = strtoul(sf, nullptr, 10); if (v >= 1 && v <= (1u << 18)) p->slab_frames_cap = (uint32_t)v; }
    *out = guard.release();
    return LBAD_OK;
}

extern "C" void lbadcu_plan_destroy(lbadcu_plan* p) {
    if (!p) return;
    DeviceScope _device_scope(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (int i = 0; i < 4; i++) if (p->scratch_done[i]) { if (p->scratch_used[i]) cudaEventSynchronize(p->scratch_done[i]); cudaEventDestroy(p->scratch_done[i]); }
    p->timer.clear(); p->timer2.clear();
    cudaFree(p->d_tw_m); cudaFree(p->d_tw_n); cudaFree(p->d_tw1); cudaFree(p->d_tw2); for (int i = 0; i < 4; i++) cudaFree(p->d_scratch[i]); for (int i = 0; i < 3; i++) cudaFree(p->d_chunk_i16[i]);
    for (int i = 0; i < 3; i++) { cudaFree(p->d_chunk_pcm[i]); cudaFree(p->d_chunk_words[i]); if (p->copy_streams[i]) cudaStreamDestroy(p->copy_streams[i]); }
    if (p->stream) cudaStreamDestroy(p->stream);
    cudaFree(p->d_score); cudaFreeHost(p->h_score);
    delete p;
}

extern "C" int lbadcu_plan_fused_supported(const lbadcu_plan* p) { return p->fused_ok ? 1 : 0; }
extern "C" int lbadcu_plan_device(const lbadcu_plan* p) { return p->device; }
extern "C" void* lbadcu_plan_stream(lbadcu_plan* p) { return p->stream; }
extern "C" uint64_t lbadcu_plan_launches(const lbadcu_plan* p) { return p->launches; }
extern "C" uint32_t lbadcu_plan_timing(lbadcu_plan* p, int which, int enable, int reset, double* total_ms) {
    LaunchTimer& t = which ? p->timer2 : p->timer;      /* 0: FFT + band-energy kernel (dominant), 1: Haar / select / pack kernel */
    uint32_t n = t.collect(total_ms, reset != 0);
    t.enabled = enable != 0;
    return n;
}

template <int B>
static int launch_haar_select(lbadcu_plan* p, const float* d_images, float* d_haar, uint32_t* d_words, uint32_t frames, cudaStream_t s) {
    const uint32_t grid = frames < (uint32_t)p->sm_count * 4 ? frames : (uint32_t)p->sm_count * 4;
    const size_t smem = 2 * (size_t)LBAD_ROWS_PER_FRAME * (B + 1) * sizeof(float);
    LBAD_CUDA_TRY(cudaFuncSetAttribute(haar_select_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    haar_select_kernel<B><<<grid, HS_THREADS, smem, s>>>(d_images, d_haar, d_words, (int)p->g.pairs, (int)p->g.words_per_plane, frames);
    p->launches++;
    LBAD_CUDA_TRY(cudaGetLastError());
    return LBAD_OK;
}

static int haar_select_dispatch(lbadcu_plan* p, const float* d_images, float* d_haar, uint32_t* d_words, uint32_t frames, cudaStream_t s) {
    switch (p->g.bands) {
        case 4:  return launch_haar_select<4>(p, d_images, d_haar, d_words, frames, s);
        case 8:  return launch_haar_select<8>(p, d_images, d_haar, d_words, frames, s);
        case 16: return launch_haar_select<16>(p, d_images, d_haar, d_words, frames, s);
        case 32: return launch_haar_select<32>(p, d_images, d_haar, d_words, frames, s);
        case 64: return launch_haar_select<64>(p, d_images, d_haar, d_words, frames, s);
    }
    return LBAD_ERR_ARG;
}

static int extract_device_slot(lbadcu_plan* p, const float* d_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride,
                               uint32_t* d_words, float* d_images, float* d_haar, int mode, void* stream, int slot);

extern "C" int lbadcu_extract_device(lbadcu_plan* p, const float* d_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride,
                                     uint32_t* d_words, float* d_images, float* d_haar, int mode, void* stream) {
    return extract_device_slot(p, d_pcm, n_clips, clip_len, clip_stride, d_words, d_images, d_haar, mode, stream, 0);
}

/* slot: which scratch buffer holds the spectral images between the two kernels — concurrent streams must not share one */
static int extract_device_slot(lbadcu_plan* p, const float* d_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride,
                               uint32_t* d_words, float* d_images, float* d_haar, int mode, void* stream, int slot) {
    if (!p || !d_pcm || !d_words) return LBAD_ERR_ARG;
    if (clip_len < p->g.window) return LBAD_ERR_ARG;
    LBAD_ON_DEVICE(p->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : p->stream;
    const uint64_t frames_per_clip = ((clip_len - p->g.window) / p->g.stride) / LBAD_ROWS_PER_FRAME;   /* m:250-255 */
    const uint64_t total_frames64 = frames_per_clip * n_clips;
    if (total_frames64 == 0) return LBAD_OK;
    if (total_frames64 > 0x7fffffffull) return LBAD_ERR_ARG;
    const uint32_t total_frames = (uint32_t)total_frames64;
    Geo g = p->g; g.frames_per_clip = (uint32_t)frames_per_clip; g.clip_stride = clip_stride;
    const bool fused = mode == 1 ? true : mode == 2 ? false : p->fused_ok;
    if (fused && !p->fused_ok) return LBAD_ERR_ARG;
    struct SlotOrder {           /* orders this call after the slot's previous user and marks the slot on every way out */
        lbadcu_plan* p; int slot; cudaStream_t s; bool active;
        ~SlotOrder() { if (active && cudaEventRecord(p->scratch_done[slot], s) == cudaSuccess) { p->scratch_stream[slot] = s; p->scratch_used[slot] = true; } }
    } slot_order{p, slot, s, d_images == nullptr};
    if (!d_images && p->scratch_used[slot] && p->scratch_stream[slot] != s) LBAD_CUDA_TRY(cudaStreamWaitEvent(s, p->scratch_done[slot], 0));
    if (fused) {
        /* fast path: FFT + bands kernel -> spectral images (global, L2-friendly 16 KB each) -> Haar/select/pack kernel */
        const bool carry = g.window == 2048 && g.stride == 64;           /* consecutive windows one pass-1 input apart: half transforms are shared */
        /* few frames (a single clip): eight CTAs per frame, so the call's latency is two windows per warp instead of sixteen */
        /* pair transposition (default for the carried kernel): twice the scratch per warp, which two CTAs per SM no longer fit with a whole
         * frame of samples each — so ONE CTA per SM runs two 8-warp groups that share the twiddle tables (LBAD_GROUPS=1: half a frame per
         * CTA iteration and two CTAs per SM instead; LBAD_TRANSPOSE=components: round 1's component-wise transposition) */
        const bool aos = carry && p->pair_transpose != 0;
        const uint32_t subs = !carry ? 1u : p->force_subs ? (uint32_t)p->force_subs : (total_frames <= (uint32_t)p->sm_count ? 8u : (aos && p->force_groups == 1) ? 2u : 1u);
        const uint32_t groups = (aos && subs == 1 && p->force_groups != 1) ? 2u : 1u;
        const uint32_t span = (LBAD_ROWS_PER_FRAME / subs - 1u) * g.stride + g.window;
        const FusedSmemLayout L = fused_layout(span, aos, groups, carry);
        /* TMA bulk copies need 16-byte aligned sources and sizes */
        bool tma_ok = ((uintptr_t)d_pcm % 16 == 0) && (clip_stride % 4 == 0) && (g.stride % 4 == 0);
        if (p->stage_mode == 0) tma_ok = false;
        auto kern = carry ? (aos ? (subs == 8 ? (p->static_range ? bands_fused_kernel<32, true, true, 8, true> : bands_fused_kernel<32, false, true, 8, true>)
                                  : subs == 2 ? (p->static_range ? bands_fused_kernel<32, true, true, 2, true> : bands_fused_kernel<32, false, true, 2, true>)
                                  : groups == 2 ? (p->static_range ? bands_fused_kernel<32, true, true, 1, true, 2> : bands_fused_kernel<32, false, true, 1, true, 2>)
                                              : (p->static_range ? bands_fused_kernel<32, true, true, 1, true> : bands_fused_kernel<32, false, true, 1, true>))
                           : subs == 8 ? (p->static_range ? bands_fused_kernel<32, true, true, 8> : bands_fused_kernel<32, false, true, 8>)
                           : subs == 2 ? (p->static_range ? bands_fused_kernel<32, true, true, 2> : bands_fused_kernel<32, false, true, 2>)
                                       : (p->static_range ? bands_fused_kernel<32, true, true> : bands_fused_kernel<32, false, true>))
                  : g.window == 2048 ? bands_fused_kernel<32, false, false> : g.window == 1024 ? bands_fused_kernel<16, false, false>
                  : g.window == 512 ? bands_fused_kernel<8, false, false> : bands_fused_kernel<4, false, false>;
        int per_sm = 0;
        for (auto& c : p->occ_cache) if (c.fn == (const void*)kern && c.smem == L.total_bytes) per_sm = c.per_sm;
        if (!per_sm) {                                                    /* first launch of this variant with this footprint */
            /* the attribute is a ceiling shared by every plan of the process: raise it to the device's limit once, never lower it */
            LBAD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_optin));
            LBAD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (int)(FUSED_THREADS * groups), L.total_bytes));
            if (per_sm < 1) per_sm = 1;
            for (auto& c : p->occ_cache) if (!c.fn) { c.fn = (const void*)kern; c.smem = L.total_bytes; c.per_sm = per_sm; break; }
        }
        const uint32_t cap = (uint32_t)(p->sm_count * per_sm);
        const uint32_t slab_cap = p->slab_frames_cap;                     /* <= 4 GB of images per slab (LBAD_SLAB_FRAMES overrides, for tests) */
        const uint32_t slab_frames = d_images ? total_frames : (total_frames < slab_cap ? total_frames : slab_cap);
        if (!d_images && p->scratch_frames[slot] < slab_frames) {
            if (p->scratch_used[slot]) LBAD_CUDA_TRY(cudaEventSynchronize(p->scratch_done[slot]));
            LBAD_CUDA_TRY(cudaStreamSynchronize(s));
            cudaFree(p->d_scratch[slot]); p->d_scratch[slot] = nullptr; p->scratch_frames[slot] = 0;
            LBAD_CUDA_TRY(cudaMalloc(&p->d_scratch[slot], (size_t)slab_frames * LBAD_ROWS_PER_FRAME * 32 * sizeof(float)));
            p->scratch_frames[slot] = slab_frames;
        }
        for (uint32_t f0 = 0; f0 < total_frames; f0 += slab_frames) {
            const uint32_t nf = total_frames - f0 < slab_frames ? total_frames - f0 : slab_frames;
            float* imgs = d_images ? d_images : p->d_scratch[slot];
            /* a slab starts at frame f0 of the flattened (clip, frame) order: hand the kernel a view that starts there */
            Geo gs = g;
            const uint32_t units = nf * subs;
            const uint32_t ctas = (units + groups - 1) / groups;
            const uint32_t grid = ctas < cap ? ctas : cap;
            p->timer.begin(s);
            kern<<<grid, FUSED_THREADS * groups, L.total_bytes, s>>>(d_pcm, imgs, p->d_tw1, p->d_tw2, gs, p->bt, L, span, units, tma_ok ? 1 : 0, f0);
            p->timer.end(s);
            p->launches++;
            LBAD_CUDA_TRY(cudaGetLastError());
            const uint32_t grid2 = nf < (uint32_t)p->sm_count * 8 ? nf : (uint32_t)p->sm_count * 8;
            p->timer2.begin(s);
            haar_select32_kernel<<<grid2, HS32_THREADS, 0, s>>>(imgs, d_haar ? d_haar + (size_t)f0 * LBAD_ROWS_PER_FRAME * 32 : nullptr,
                                                                d_words + (size_t)f0 * 2 * g.words_per_plane, (int)g.pairs, (int)g.words_per_plane, nf);
            p->timer2.end(s);
            p->launches++;
            LBAD_CUDA_TRY(cudaGetLastError());
        }
        return LBAD_OK;
    }
    /* generic: bands -> (scratch) images -> Haar/select, in slabs of frames when no dump buffer was given */
    const uint32_t M = g.window / 2;
    const size_t smem = (size_t)GEN_WARPS * M * (sizeof(float2) + sizeof(float));
    LBAD_CUDA_TRY(cudaFuncSetAttribute(bands_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t clips_per_slab = n_clips;
    if (!d_images) {
        const uint64_t want_clips = 8192 / frames_per_clip ? 8192 / frames_per_clip : 1;      /* ~8192 frames (128 MB at B=32) per slab */
        clips_per_slab = (uint32_t)(want_clips < n_clips ? want_clips : n_clips);
        const size_t need = (size_t)clips_per_slab * frames_per_clip;
        if (p->scratch_frames[slot] < need * 2) {                         /* sized for 64 bands: twice the 32-band frame size */
            if (p->scratch_used[slot]) LBAD_CUDA_TRY(cudaEventSynchronize(p->scratch_done[slot]));
            LBAD_CUDA_TRY(cudaStreamSynchronize(s));
            cudaFree(p->d_scratch[slot]); p->d_scratch[slot] = nullptr; p->scratch_frames[slot] = 0;
            LBAD_CUDA_TRY(cudaMalloc(&p->d_scratch[slot], need * 2 * LBAD_ROWS_PER_FRAME * 32 * sizeof(float)));
            p->scratch_frames[slot] = need * 2;
        }
    }
    for (uint32_t c0 = 0; c0 < n_clips; c0 += clips_per_slab) {
        const uint32_t nc = (n_clips - c0) < clips_per_slab ? (n_clips - c0) : clips_per_slab;
        const uint32_t nf = (uint32_t)(nc * frames_per_clip);
        const uint64_t nw = (uint64_t)nf * LBAD_ROWS_PER_FRAME;
        float* imgs = d_images ? d_images + (size_t)c0 * frames_per_clip * LBAD_ROWS_PER_FRAME * g.bands : p->d_scratch[slot];
        const uint64_t want = (nw + GEN_WARPS - 1) / GEN_WARPS;
        const uint32_t grid = (uint32_t)(want < (uint64_t)p->sm_count * 16 ? want : (uint64_t)p->sm_count * 16);
        p->timer.begin(s);
        bands_generic_kernel<<<grid, GEN_WARPS * 32, smem, s>>>(d_pcm + (uint64_t)c0 * clip_stride, imgs, p->d_tw_m, p->d_tw_n, g, p->bt, nw);
        p->timer.end(s);
        p->launches++;
        LBAD_CUDA_TRY(cudaGetLastError());
        float* haar = d_haar ? d_haar + (size_t)c0 * frames_per_clip * LBAD_ROWS_PER_FRAME * g.bands : nullptr;
        int e = haar_select_dispatch(p, imgs, haar, d_words + (size_t)c0 * frames_per_clip * 2 * g.words_per_plane, nf, s);
        if (e != LBAD_OK) return e;
    }
    return LBAD_OK;
}

extern "C" int lbadcu_transform_images_host(lbadcu_plan* p, const float* h_images, uint32_t count, float* h_haar, uint32_t* h_words) {
    if (!p || !h_images || count == 0) return LBAD_ERR_ARG;
    LBAD_ON_DEVICE(p->device);
    const size_t n = (size_t)count * LBAD_ROWS_PER_FRAME * p->g.bands, nw = (size_t)count * 2 * p->g.words_per_plane;
    DevBuf<float> d_img, d_haar; DevBuf<uint32_t> d_words;
    LBAD_CUDA_TRY(d_img.alloc(n)); LBAD_CUDA_TRY(d_haar.alloc(n)); LBAD_CUDA_TRY(d_words.alloc(nw));
    LBAD_CUDA_TRY(cudaMemcpyAsync(d_img, h_images, n * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    /* the kernel the extraction path itself would run on these images: the register kernel for 32 bands (LBAD_TRANSFORM=generic forces the other) */
    int e = LBAD_OK;
    if (p->g.bands == 32 && !p->transform_generic) {
        const uint32_t grid2 = count < (uint32_t)p->sm_count * 8 ? count : (uint32_t)p->sm_count * 8;
        haar_select32_kernel<<<grid2, HS32_THREADS, 0, p->stream>>>(d_img, d_haar, d_words, (int)p->g.pairs, (int)p->g.words_per_plane, count);
        p->launches++;
        LBAD_CUDA_TRY(cudaGetLastError());
    } else e = haar_select_dispatch(p, d_img, d_haar, d_words, count, p->stream);
    if (e == LBAD_OK) {
        if (h_haar) LBAD_CUDA_TRY(cudaMemcpyAsync(h_haar, d_haar, n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
        if (h_words) LBAD_CUDA_TRY(cudaMemcpyAsync(h_words, d_words, nw * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    }
    LBAD_CUDA_TRY(cudaStreamSynchronize(p->stream));
    return e;
}

__global__ void i16_to_f32_kernel(const int16_t* __restrict__ in, float* __restrict__ out, const size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (float)in[i] * (1.0f / 32768.0f);
}

/* Host-memory front end.  Clips are processed in chunks on three streams so that the H2D copy of chunk i+1, the
 * kernels of chunk i and the D2H of chunk i-1 overlap (they do when the caller's buffers are pinned).  Each chunk
 * buffer has its own spectral-image scratch (slots 1..3), because the streams run concurrently.
 * sample_bytes: 4 = float32 PCM, 2 = signed 16-bit PCM (converted on the device as x / 32768, which is exact). */
/* cursor (optional): a clip counter SHARED by several plans working on the same batch from different host threads (one plan per GPU,
 * lbadcu_extract_host_shared): a plan takes its next chunk from it only when the buffer that chunk goes to has been drained, so every
 * GPU keeps three chunks in flight and the batch divides itself by how fast each GPU gets its data — the host links of one box differ. */
static int extract_host_impl(lbadcu_plan* p, const void* h_pcm_v, int sample_bytes, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride,
                             uint32_t* h_words, float* h_images, float* h_haar, int mode, uint64_t* cursor = nullptr) {
    if (!p || !h_pcm_v || !h_words || n_clips == 0) return LBAD_ERR_ARG;
    if (clip_len < p->g.window) return LBAD_ERR_ARG;
    LBAD_ON_DEVICE(p->device);
    const char* h_pcm = static_cast<const char*>(h_pcm_v);
    const uint64_t frames_per_clip = ((clip_len - p->g.window) / p->g.stride) / LBAD_ROWS_PER_FRAME;
    if (frames_per_clip == 0) return LBAD_OK;
    const size_t words_per_clip = (size_t)frames_per_clip * 2 * p->g.words_per_plane;
    const size_t img_per_clip = (size_t)frames_per_clip * LBAD_ROWS_PER_FRAME * p->g.bands;
    const uint64_t clip_pad = (clip_len + 7) & ~7ull;                  /* device clip stride: keeps every clip 16-byte aligned in both sample formats */
    uint64_t clips_per_chunk = (48ull << 20) / clip_pad;                /* ~192 MB of float PCM per chunk */
    if (p->chunk_clips_override) clips_per_chunk = p->chunk_clips_override;
    if (clips_per_chunk < 1) clips_per_chunk = 1;
    if (clips_per_chunk > n_clips) clips_per_chunk = n_clips;
    if (cursor) {                                                       /* plans that share a batch must cut it alike: the first one to arrive fixes the chunk size (cursor[1]) */
        uint64_t unset = 0;
        __atomic_compare_exchange_n(&cursor[1], &unset, clips_per_chunk, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE);
        clips_per_chunk = __atomic_load_n(&cursor[1], __ATOMIC_ACQUIRE);
    }
    const size_t need_pcm = (size_t)clips_per_chunk * clip_pad, need_words = (size_t)clips_per_chunk * words_per_clip;
    const int nbuf = n_clips > clips_per_chunk ? 3 : 1;
    if (p->chunk_pcm_floats < need_pcm || p->chunk_words < need_words || (sample_bytes == 2 && p->chunk_i16 < need_pcm)) {
        for (int i = 0; i < 3; i++) {
            cudaFree(p->d_chunk_pcm[i]); cudaFree(p->d_chunk_words[i]); cudaFree(p->d_chunk_i16[i]);
            p->d_chunk_pcm[i] = nullptr; p->d_chunk_words[i] = nullptr; p->d_chunk_i16[i] = nullptr;
        }
        p->chunk_pcm_floats = p->chunk_words = p->chunk_i16 = 0;
    }
    for (int i = 0; i < nbuf; i++) {
        if (!p->d_chunk_pcm[i]) LBAD_CUDA_TRY(cudaMalloc(&p->d_chunk_pcm[i], need_pcm * sizeof(float)));
        if (!p->d_chunk_words[i]) LBAD_CUDA_TRY(cudaMalloc(&p->d_chunk_words[i], need_words * sizeof(uint32_t)));
        if (sample_bytes == 2 && !p->d_chunk_i16[i]) LBAD_CUDA_TRY(cudaMalloc(&p->d_chunk_i16[i], need_pcm * sizeof(int16_t)));
    }
    p->chunk_pcm_floats = need_pcm; p->chunk_words = need_words; if (sample_bytes == 2) p->chunk_i16 = need_pcm;
    DevBuf<float> d_img, d_haar;
    if (h_images) LBAD_CUDA_TRY(d_img.alloc((size_t)clips_per_chunk * img_per_clip));
    if (h_haar) LBAD_CUDA_TRY(d_haar.alloc((size_t)clips_per_chunk * img_per_clip));
    int rc = LBAD_OK;
    uint32_t chunk = 0;
    for (uint64_t own = 0;; own += clips_per_chunk, chunk++) {
        const int b = (int)(chunk % (uint32_t)nbuf);
        cudaStream_t s = nbuf == 1 ? p->stream : p->copy_streams[b];
        if (cursor && chunk >= (uint32_t)nbuf) LBAD_CUDA_TRY(cudaStreamSynchronize(s));      /* shared batch: another chunk only once this buffer's last one is through */
        const uint64_t c0 = cursor ? __atomic_fetch_add(&cursor[0], clips_per_chunk, __ATOMIC_RELAXED) : own;
        if (c0 >= n_clips) break;
        const uint32_t nc = (uint32_t)((n_clips - c0) < clips_per_chunk ? (n_clips - c0) : clips_per_chunk);
        void* d_in = sample_bytes == 2 ? static_cast<void*>(p->d_chunk_i16[b]) : static_cast<void*>(p->d_chunk_pcm[b]);
        const char* src = h_pcm + c0 * clip_stride * sample_bytes;
        /* one contiguous copy when the host layout already has the device stride — up to the last clip's last SAMPLE, not its padded end:
         * the caller's buffer may end there */
        if (clip_stride == clip_pad) LBAD_CUDA_TRY(cudaMemcpyAsync(d_in, src, ((size_t)(nc - 1) * clip_pad + clip_len) * sample_bytes, cudaMemcpyHostToDevice, s));
        else LBAD_CUDA_TRY(cudaMemcpy2DAsync(d_in, clip_pad * sample_bytes, src, clip_stride * sample_bytes, clip_len * sample_bytes, nc, cudaMemcpyHostToDevice, s));
        if (sample_bytes == 2) {
            i16_to_f32_kernel<<<p->sm_count * 8, 256, 0, s>>>(p->d_chunk_i16[b], p->d_chunk_pcm[b], (size_t)nc * clip_pad);
            p->launches++;
            LBAD_CUDA_TRY(cudaGetLastError());
        }
        rc = extract_device_slot(p, p->d_chunk_pcm[b], nc, clip_len, clip_pad, p->d_chunk_words[b], d_img, d_haar, mode, s, 1 + b);
        if (rc != LBAD_OK) break;
        LBAD_CUDA_TRY(cudaMemcpyAsync(h_words + c0 * words_per_clip, p->d_chunk_words[b], (size_t)nc * words_per_clip * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (h_images) LBAD_CUDA_TRY(cudaMemcpyAsync(h_images + c0 * img_per_clip, d_img, (size_t)nc * img_per_clip * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (h_haar) LBAD_CUDA_TRY(cudaMemcpyAsync(h_haar + c0 * img_per_clip, d_haar, (size_t)nc * img_per_clip * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (h_images || h_haar) LBAD_CUDA_TRY(cudaStreamSynchronize(s));      /* the dump buffers are shared between chunks */
    }
    if (nbuf > 1) { for (int i = 0; i < 3; i++) LBAD_CUDA_TRY(cudaStreamSynchronize(p->copy_streams[i])); }      /* one chunk: everything ran on the plan's stream */
    LBAD_CUDA_TRY(cudaStreamSynchronize(p->stream));
    return rc;
}

extern "C" int lbadcu_extract_host(lbadcu_plan* p, const float* h_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride,
                                   uint32_t* h_words, float* h_images, float* h_haar, int mode) {
    return extract_host_impl(p, h_pcm, 4, n_clips, clip_len, clip_stride, h_words, h_images, h_haar, mode);
}

/* the same batch shared between several plans (see extract_host_impl): every caller passes the whole batch and the same cursor,
 * two zero-initialised 64-bit words (clips handed out so far; clips per chunk, fixed by the first plan to arrive) */
extern "C" int lbadcu_extract_host_shared(lbadcu_plan* p, const float* h_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride, uint32_t* h_words, uint64_t* cursor) {
    if (!cursor) return LBAD_ERR_ARG;
    return extract_host_impl(p, h_pcm, 4, n_clips, clip_len, clip_stride, h_words, nullptr, nullptr, 0, cursor);
}

extern "C" int lbadcu_extract_host_i16(lbadcu_plan* p, const int16_t* h_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride, uint32_t* h_words) {
    return extract_host_impl(p, h_pcm, 2, n_clips, clip_len, clip_stride, h_words, nullptr, nullptr, 0);
}

/* LBAudioDetectiveCompareAudioURLs (m:442-464) without a host round trip between its three steps: both clips go up on the plan's
 * stream, are fingerprinted there — as one two-clip batch when the lengths agree, so the two clips' frames run side by side — and
 * compare_pair_kernel reads the packed words where the extraction left them.  One synchronisation, four bytes come back. */
extern "C" int lbadcu_compare_pcm_host(lbadcu_plan* p, const float* h1, uint64_t n1, const float* h2, uint64_t n2, uint32_t pairs, float* out) {
    if (!p || !h1 || !h2 || !out) return LBAD_ERR_ARG;
    if (n1 < p->g.window || n2 < p->g.window) return LBAD_ERR_ARG;
    LBAD_ON_DEVICE(p->device);
    const uint64_t f1 = ((n1 - p->g.window) / p->g.stride) / LBAD_ROWS_PER_FRAME, f2 = ((n2 - p->g.window) / p->g.stride) / LBAD_ROWS_PER_FRAME;   /* m:250-255 */
    if (f1 == 0 || f2 == 0 || f1 + f2 > 0x7fffffffull) return LBAD_ERR_ARG;
    const uint64_t pad1 = (n1 + 7) & ~7ull, pad2 = (n2 + 7) & ~7ull;
    const size_t W2 = 2 * (size_t)p->g.words_per_plane, need_pcm = (size_t)(pad1 + pad2), need_words = (size_t)(f1 + f2) * W2;
    if (p->chunk_pcm_floats < need_pcm || p->chunk_words < need_words || !p->d_chunk_pcm[0] || !p->d_chunk_words[0]) {
        for (int i = 0; i < 3; i++) {                                    /* the chunk buffers are shared with the batch front end: same sizes for all three */
            LBAD_CUDA_TRY(cudaStreamSynchronize(p->copy_streams[i]));
            cudaFree(p->d_chunk_pcm[i]); cudaFree(p->d_chunk_words[i]); cudaFree(p->d_chunk_i16[i]);
            p->d_chunk_pcm[i] = nullptr; p->d_chunk_words[i] = nullptr; p->d_chunk_i16[i] = nullptr;
        }
        LBAD_CUDA_TRY(cudaStreamSynchronize(p->stream));
        p->chunk_pcm_floats = p->chunk_words = p->chunk_i16 = 0;
        LBAD_CUDA_TRY(cudaMalloc(&p->d_chunk_pcm[0], need_pcm * sizeof(float)));
        LBAD_CUDA_TRY(cudaMalloc(&p->d_chunk_words[0], need_words * sizeof(uint32_t)));
        p->chunk_pcm_floats = need_pcm; p->chunk_words = need_words;
    }
    if (!p->d_score) { LBAD_CUDA_TRY(cudaMalloc(&p->d_score, sizeof(float))); LBAD_CUDA_TRY(cudaHostAlloc(&p->h_score, sizeof(float), cudaHostAllocDefault)); }
    cudaStream_t s = p->stream;
    float* d1 = p->d_chunk_pcm[0]; float* d2 = d1 + pad1;
    uint32_t* w1 = p->d_chunk_words[0]; uint32_t* w2 = w1 + f1 * W2;
    LBAD_CUDA_TRY(cudaMemcpyAsync(d1, h1, n1 * sizeof(float), cudaMemcpyHostToDevice, s));
    LBAD_CUDA_TRY(cudaMemcpyAsync(d2, h2, n2 * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc;
    if (n1 == n2) rc = extract_device_slot(p, d1, 2, n1, pad1, w1, nullptr, nullptr, 0, s, 1);
    else {
        rc = extract_device_slot(p, d1, 1, n1, pad1, w1, nullptr, nullptr, 0, s, 1);
        if (rc == LBAD_OK) rc = extract_device_slot(p, d2, 1, n2, pad2, w2, nullptr, nullptr, 0, s, 1);
    }
    if (rc == LBAD_OK) { rc = lbadcu_compare_pair_device(p->g.words_per_plane, pairs, w1, (uint32_t)f1, w2, (uint32_t)f2, p->d_score, s); p->launches++; }
    if (rc != LBAD_OK) { cudaStreamSynchronize(s); return rc; }
    LBAD_CUDA_TRY(cudaMemcpyAsync(p->h_score, p->d_score, sizeof(float), cudaMemcpyDeviceToHost, s));
    LBAD_CUDA_TRY(cudaStreamSynchronize(s));
    *out = *p->h_score;
    return LBAD_OK;
}
