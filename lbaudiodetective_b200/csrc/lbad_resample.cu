/*
 * lbad_resample.cu — recording-rate -> processing-rate conversion on the device (sm_100a), the step ExtAudioFile's client
 * format performs in the reference (LBAudioDetective.m:229, m:275; AudioToolbox itself is out of scope).  The conversion is
 * the two-stage band-limited decimator defined in include/LBAudioDetectiveResample.h; the tables are designed on the host
 * (lbad_resample_design.c) and this file evaluates them: one CTA (128 threads) per tile of 128 or 256 output samples of one clip, the input span of
 * the tile staged once in shared memory (every recorded sample is read from HBM once per tile it touches; neighbouring tiles
 * overlap by the filter length, served by L2), stage 1 into a second shared tile, stage 2 from there.  Sums are float32 FMA
 * chains in increasing tap order, so the result equals oracle/lbad_oracle.c bit for bit.
 */
#include "lbad_common.cuh"
#include <math.h>
#include <type_traits>

namespace lbad {

constexpr int RS_THREADS = 128;              /* a tile is RS_THREADS x OPT output samples (OPT = outputs per thread in stage 2) */
constexpr int RS_PHASES = LBAD_RS_PHASES;

struct ResampleParams {
    uint32_t D, H1, T1, H2, T2;
    double rho2;
    uint64_t in_len, in_stride, out_len, out_stride;
    uint32_t tiles_per_clip, ny_max, nx_max, tile;      /* tile: output samples per CTA iteration */
    uint32_t nip, xs_floats;         /* D = 4: floats per polyphase plane of the staged input; floats reserved for the staged input */
};

/* The stage-1 taps for D = 4 depend on nothing but D (sinc((t - 24) / 4) / 4 under a Kaiser window), so one constant copy serves every
 * resampler of the process; as constant-bank operands of the FMAs they cost neither registers nor loads. */
__constant__ float c_g_d4[49];

/* D4: integer decimation by 4 with the 49-tap stage-1 filter — the 44.1 kHz case.  The tile's input starts on a multiple of 4
 * samples, so thread j reads its 49 inputs as twelve conflict-free 128-bit shared loads plus one scalar.  T2C: the number of
 * stage-2 taps when known at compile time (42 for 44.1 kHz -> 5512 Hz), 0 = run-time.  vec_ok: the clips start on 16-byte
 * boundaries, so interior tiles are staged with 128-bit loads.  OPT: outputs per thread in stage 2.  OPT = 2 (rho2 in [2, 3), T2C
 * given) pairs outputs m, m + 1: while consecutive outputs sit exactly two stage-1 samples apart — all but one step in 1 / frac(rho2)
 * — a thread reads the 44 samples its two outputs share as twelve 128-bit loads and, while both outputs use the same coarse phase,
 * one set of coefficient rows: 35 instead of 64 shared-memory wavefronts per 32 outputs (stage 2 was what bound the kernel).  The
 * FMA chain of every output is the one of the single-output form, so the result does not change by a bit. */
/* Where 128-bit group I of a polyphase plane is kept: a thread's eight stage-1 outputs start two groups after its neighbour's, so the
 * eight lanes of a 128-bit access phase read groups s, s + 2, .., s + 14 — every bank twice.  Exchanging odd and even groups in every
 * second block of eight puts the groups that are eight apart on different banks: conflict-free for any s. */
__device__ __forceinline__ uint32_t plane_quad(const uint32_t I) { return I ^ ((I >> 3) & 1u); }
__device__ __forceinline__ uint32_t plane_pos(const uint32_t i) { return i ^ (((i >> 5) & 1u) << 2); }      /* the same for float index i = 4 I + c */

template <bool D4, int T2C, int OPT>
__global__ void __launch_bounds__(RS_THREADS)
resample_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ g, const float* __restrict__ hc, const ResampleParams P,
                const uint64_t total_tiles, const int vec_ok) {
    extern __shared__ __align__(16) float rs_smem[];
    float* xs = rs_smem;                                       /* [nx_max], 16-byte aligned */
    float* y1s = xs + P.xs_floats;                             /* [ny_max + 4] */
    float* gs = y1s + ((P.ny_max + 8 + 3) & ~3u);              /* [T1] */
    float* hcs = gs + ((P.T1 + 3) & ~3u);                      /* [PHASES + 1][T2P], rows padded to T2P = a multiple of 4 floats (zeros) */
    const uint32_t T2 = T2C ? (uint32_t)T2C : P.T2;
    const uint32_t T2P = (T2 + 3) & ~3u;
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < P.T1; i += RS_THREADS) gs[i] = g[i];
    for (uint32_t i = tid; i < (RS_PHASES + 1) * T2P; i += RS_THREADS) { const uint32_t r = i / T2P, c = i % T2P; hcs[i] = c < T2 ? __ldg(hc + r * T2 + c) : 0.0f; }
    for (uint32_t tile = blockIdx.x; tile < (uint32_t)total_tiles; tile += gridDim.x) {      /* the host keeps total_tiles and out_len below 2^31 */
        const uint32_t clip = tile / P.tiles_per_clip;
        constexpr uint32_t TILE = RS_THREADS * OPT;
        const uint32_t m0 = (tile % P.tiles_per_clip) * TILE;
        const uint32_t m1 = (m0 + TILE < (uint32_t)P.out_len) ? m0 + TILE : (uint32_t)P.out_len;
        const int64_t i0_lo = (int64_t)floor((double)m0 * P.rho2), i0_hi = (int64_t)floor((double)(m1 - 1) * P.rho2);
        const int64_t y_lo = i0_lo - (int64_t)P.H2 + 1, y_hi = i0_hi + (int64_t)P.H2;
        const uint32_t ny = (uint32_t)(y_hi - y_lo + 1);
        const float* src = in + (uint64_t)clip * P.in_stride;
        __syncthreads();                                        /* the previous tile's readers are done with the shared tiles (first pass: the tables are in place) */
        if (P.D > 1) {
            const int64_t x_lo = (int64_t)P.D * y_lo - (int64_t)P.H1;
            const uint32_t nx = P.D * (ny - 1) + P.T1;
            if constexpr (D4) {
                /* the input tile is staged de-interleaved into its four polyphase planes xp[p][i] = x[x_lo + 4 i + p] (x_lo is a multiple
                 * of 4): a thread then produces EIGHT consecutive stage-1 outputs from five 128-bit loads per plane, 20 loads for 392 FMAs,
                 * instead of 13 loads per output — stage 1 was bound by shared-memory wavefronts */
                const uint32_t ni = (nx + 3) / 4, nip = P.nip;                              /* plane length used / allocated (multiple of 4 floats, >= ny_max + 24) */
                if (vec_ok && x_lo >= 0 && (uint64_t)x_lo + 4ull * ni <= P.in_len) {
                    const float4* s4 = reinterpret_cast<const float4*>(src + x_lo);
                    for (uint32_t i = tid; i < ni; i += RS_THREADS) { const float4 v = __ldg(s4 + i); const uint32_t j = plane_pos(i); xs[j] = v.x; xs[nip + j] = v.y; xs[2 * nip + j] = v.z; xs[3 * nip + j] = v.w; }
                } else {
                    for (uint32_t i = tid; i < 4 * ni; i += RS_THREADS) { const int64_t k = x_lo + i; xs[(i & 3) * nip + plane_pos(i >> 2)] = (k >= 0 && (uint64_t)k < P.in_len) ? __ldg(src + k) : 0.0f; }
                }
                __syncthreads();
                for (uint32_t j0 = 8 * tid; j0 < ny; j0 += 8 * RS_THREADS) {             /* stage 1: outputs j0 .. j0 + 7 */
                    float acc[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                    for (int pl = 0; pl < 4; pl++) {
                        const float4* x4 = reinterpret_cast<const float4*>(xs + pl * nip);
                        float v[20];
#pragma unroll
                        for (int q = 0; q < 5; q++) { const float4 t = x4[plane_quad((j0 >> 2) + q)]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
#pragma unroll
                        for (int m = 0; m < (pl == 0 ? 13 : 12); m++) {                   /* tap t = 4 m + pl multiplies plane sample (output + m) */
                            const float c = c_g_d4[4 * m + pl];
#pragma unroll
                            for (int o = 0; o < 8; o++) acc[o] = fmaf(c, v[m + o], acc[o]);
                        }
                    }
                    *reinterpret_cast<float4*>(y1s + j0) = make_float4(acc[0], acc[1], acc[2], acc[3]);      /* y1s has room for the up to seven outputs past ny */
                    *reinterpret_cast<float4*>(y1s + j0 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                }
            } else {
                for (uint32_t i = tid; i < nx; i += RS_THREADS) { const int64_t k = x_lo + i; xs[i] = (k >= 0 && (uint64_t)k < P.in_len) ? __ldg(src + k) : 0.0f; }
                __syncthreads();
                for (uint32_t j = tid; j < ny; j += RS_THREADS) {   /* stage 1, taps in polyphase order: p outer, t = D m + p */
                    float acc = 0.0f;
                    const float* x = xs + (size_t)P.D * j;
                    for (uint32_t pl = 0; pl < P.D; pl++) for (uint32_t t = pl; t < P.T1; t += P.D) acc = fmaf(gs[t], x[t], acc);
                    y1s[j] = acc;
                }
            }
        } else {
            for (uint32_t i = tid; i < ny; i += RS_THREADS) { const int64_t k = y_lo + i; y1s[i] = (k >= 0 && (uint64_t)k < P.in_len) ? __ldg(src + k) : 0.0f; }
        }
        __syncthreads();
        /* stage 2 for one output: position in the stage-1 sequence, coarse phase, interpolation weight */
        auto single_output = [&](const uint32_t m) {
            const double pos = (double)m * P.rho2;
            const double fl = floor(pos);
            const double fp = (pos - fl) * (double)RS_PHASES;
            const int p = (int)fp;
            const float a = (float)(fp - (double)p);
            const int yb = (int)((int64_t)fl - (int64_t)P.H2 + 1 - y_lo);   /* my first stage-1 sample inside the tile */
            const float4* h0 = reinterpret_cast<const float4*>(hcs + (size_t)p * T2P);      /* warp-uniform except across a phase step: broadcast loads */
            const float4* h1 = reinterpret_cast<const float4*>(hcs + (size_t)(p + 1) * T2P);
            float s0 = 0.0f, s1 = 0.0f;
            /* samples by 64-bit loads from the even index at or below yb; `odd` shifts the window by one (same FMA order either way) */
            const float2* y2 = reinterpret_cast<const float2*>(y1s + (yb & ~1));
            const bool odd = (yb & 1) != 0;
            auto taps4 = [&](const uint32_t q, const float v0, const float v1, const float v2, const float v3) {
                const float4 c0 = h0[q], c1 = h1[q];
                const uint32_t i = 4 * q;                       /* the padded taps (i >= T2) have zero coefficients but must not enter the chain */
                s0 = fmaf(c0.x, v0, s0); s1 = fmaf(c1.x, v0, s1);
                if (i + 1 < T2) { s0 = fmaf(c0.y, v1, s0); s1 = fmaf(c1.y, v1, s1); }
                if (i + 2 < T2) { s0 = fmaf(c0.z, v2, s0); s1 = fmaf(c1.z, v2, s1); }
                if (i + 3 < T2) { s0 = fmaf(c0.w, v3, s0); s1 = fmaf(c1.w, v3, s1); }
            };
            const uint32_t quads = T2C ? (uint32_t)(T2C + 3) / 4 : T2P / 4;
            if (odd) {                                          /* (almost always uniform over the warp: the parity changes once per 1 / frac(rho2) outputs) */
                float carry = y2[0].y;
#pragma unroll
                for (uint32_t q = 0; q < quads; q++) { const float2 a2 = y2[2 * q + 1], b2 = y2[2 * q + 2]; taps4(q, carry, a2.x, a2.y, b2.x); carry = b2.y; }
            } else {
#pragma unroll
                for (uint32_t q = 0; q < quads; q++) { const float2 a2 = y2[2 * q], b2 = y2[2 * q + 1]; taps4(q, a2.x, a2.y, b2.x, b2.y); }
            }
            out[(uint64_t)clip * P.out_stride + m] = fmaf(a, s1 - s0, s0);
        };
        if constexpr (OPT == 1) {
            if (m0 + tid < m1) single_output(m0 + tid);
        } else {
            static_assert(OPT == 2 && T2C > 0, "the paired form needs the tap count at compile time");
            const int lane = tid & 31;
            const uint32_t ma = m0 + 2 * tid, mb = ma + 1;
            const double pos_a = (double)ma * P.rho2, pos_b = (double)mb * P.rho2;
            const double fl_a = floor(pos_a), fl_b = floor(pos_b);
            const double fp_a = (pos_a - fl_a) * (double)RS_PHASES, fp_b = (pos_b - fl_b) * (double)RS_PHASES;
            const int p_a = (int)fp_a, p_b = (int)fp_b;
            const int yb = (int)((int64_t)fl_a - (int64_t)P.H2 + 1 - y_lo), yb_b = (int)((int64_t)fl_b - (int64_t)P.H2 + 1 - y_lo);
            const int yb0 = __shfl_sync(0xffffffffu, yb, 0);
            /* the paired form: every output of the warp exists and every step inside the warp's 64 outputs is exactly two samples */
            if (__all_sync(0xffffffffu, mb < m1 && yb == yb0 + 4 * lane && yb_b == yb + 2)) {
                const bool same = __all_sync(0xffffffffu, p_a == p_b) != 0;              /* one set of coefficient rows serves both outputs */
                const float4* y4 = reinterpret_cast<const float4*>(y1s + (yb & ~3));
                float w[48];                                    /* stage-1 samples (yb & ~3) .. + 47; taps use w[r .. r + 43], r = yb & 3 (warp-uniform) */
#pragma unroll
                for (int q = 0; q < 12; q++) { const float4 v = y4[q]; w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
                const float4* h0a = reinterpret_cast<const float4*>(hcs + (size_t)p_a * T2P);
                const float4* h1a = reinterpret_cast<const float4*>(hcs + (size_t)(p_a + 1) * T2P);
                const float4* h0b = reinterpret_cast<const float4*>(hcs + (size_t)p_b * T2P);
                const float4* h1b = reinterpret_cast<const float4*>(hcs + (size_t)(p_b + 1) * T2P);
                float sa0 = 0.0f, sa1 = 0.0f, sb0 = 0.0f, sb1 = 0.0f;
                auto chains = [&](auto rc, auto sc) {
                    constexpr int r = decltype(rc)::value;
                    constexpr bool SAME = decltype(sc)::value;  /* both outputs on one coarse phase: b reads a's coefficient registers */
#pragma unroll
                    for (int q = 0; q < (T2C + 3) / 4; q++) {
                        const float4 ca0 = h0a[q], ca1 = h1a[q];
                        const float4 cb0 = SAME ? ca0 : h0b[q], cb1 = SAME ? ca1 : h1b[q];
                        const float a0[4] = {ca0.x, ca0.y, ca0.z, ca0.w}, a1[4] = {ca1.x, ca1.y, ca1.z, ca1.w};
                        const float b0[4] = {cb0.x, cb0.y, cb0.z, cb0.w}, b1[4] = {cb1.x, cb1.y, cb1.z, cb1.w};
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int i = 4 * q + j;
                            if (i < T2C) {                      /* the padded taps have zero coefficients but must not enter the chain */
                                sa0 = fmaf(a0[j], w[r + i], sa0); sa1 = fmaf(a1[j], w[r + i], sa1);
                                sb0 = fmaf(b0[j], w[r + 2 + i], sb0); sb1 = fmaf(b1[j], w[r + 2 + i], sb1);
                            }
                        }
                    }
                };
                auto by_alignment = [&](auto sc) {
                    switch (yb0 & 3) {                          /* warp-uniform */
                        case 0: chains(std::integral_constant<int, 0>{}, sc); break;
                        case 1: chains(std::integral_constant<int, 1>{}, sc); break;
                        case 2: chains(std::integral_constant<int, 2>{}, sc); break;
                        default: chains(std::integral_constant<int, 3>{}, sc); break;
                    }
                };
                if (same) by_alignment(std::true_type{}); else by_alignment(std::false_type{});
                float* o = out + (uint64_t)clip * P.out_stride + ma;
                o[0] = fmaf((float)(fp_a - (double)p_a), sa1 - sa0, sa0);
                o[1] = fmaf((float)(fp_b - (double)p_b), sb1 - sb0, sb0);
            } else {
                if (ma < m1) single_output(ma);
                if (mb < m1) single_output(mb);
            }
        }
    }
}

}  // namespace lbad

using namespace lbad;

struct lbadcu_resampler {
    ResampleParams P;
    float *d_g = nullptr, *d_hc = nullptr;
    int device = 0, sm_count = 0;
    size_t smem_base = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
};

extern "C" int lbadcu_resampler_create(const lbadcu_resample_design* d, lbadcu_resampler** out) {
    *out = nullptr;
    if (lbadcu_device_available() != LBAD_OK) { set_error("no CUDA device available (this library has no CPU fallback)"); return LBAD_ERR_NODEVICE; }
    if (!d || !d->hc || d->T2 == 0 || (d->D > 1 && !d->g)) return LBAD_ERR_ARG;
    lbadcu_resampler* r = new lbadcu_resampler();
    Guard<lbadcu_resampler> guard(r, lbadcu_resampler_destroy);
    LBAD_CUDA_TRY(cudaGetDevice(&r->device));
    cudaDeviceProp prop; LBAD_CUDA_TRY(cudaGetDeviceProperties(&prop, r->device));
    r->sm_count = prop.multiProcessorCount;
    ResampleParams& P = r->P;
    P.D = d->D; P.H1 = d->H1; P.T1 = d->D > 1 ? d->T1 : 0; P.H2 = d->H2; P.T2 = d->T2; P.rho2 = d->rho2;
    /* two outputs per thread in stage 2 (tiles of 512) where the paired form applies: the 49-tap D = 4 first stage, 42 taps, rho2 in [2, 3) */
    P.tile = (d->D == 4 && d->T1 == 49 && d->T2 == 42 && d->rho2 >= 2.0 && d->rho2 < 3.0) ? 2 * RS_THREADS : RS_THREADS;
    P.ny_max = (uint32_t)floor(P.tile * d->rho2) + 2 * d->H2 + 2;
    P.nx_max = d->D > 1 ? d->D * (P.ny_max - 1) + d->T1 : 0;
    P.nip = (P.nx_max / 4 + 12 + 3) & ~3u;          /* >= ny_max + 24: the five 128-bit loads of the last thread's eight outputs stay inside the plane */
    P.xs_floats = (d->D == 4 && d->T1 == 49) ? 4 * P.nip : ((P.nx_max + 3) & ~3u);
    const size_t nhc = (size_t)(RS_PHASES + 1) * d->T2;
    /* shared layout: input tile, stage-1 tile (+8: the 64- and 128-bit sample loads of stage 2 may read up to seven floats past it), taps, padded coarse-phase rows */
    r->smem_base = (size_t)P.xs_floats + (((size_t)P.ny_max + 8 + 3) & ~3u) + (((size_t)P.T1 + 3) & ~3u) + (size_t)(RS_PHASES + 1) * ((d->T2 + 3) & ~3u);
    if (r->smem_base * sizeof(float) > prop.sharedMemPerBlockOptin) { set_error("resampler: the filters for this rate pair do not fit in shared memory"); return LBAD_ERR_ARG; }
    LBAD_CUDA_TRY(cudaMalloc(&r->d_hc, nhc * sizeof(float)));
    LBAD_CUDA_TRY(cudaMemcpy(r->d_hc, d->hc, nhc * sizeof(float), cudaMemcpyHostToDevice));
    LBAD_CUDA_TRY(cudaMalloc(&r->d_g, (P.T1 ? P.T1 : 1) * sizeof(float)));
    if (P.T1) LBAD_CUDA_TRY(cudaMemcpy(r->d_g, d->g, P.T1 * sizeof(float), cudaMemcpyHostToDevice));
    if (P.D == 4 && P.T1 == 49) LBAD_CUDA_TRY(cudaMemcpyToSymbol(c_g_d4, d->g, 49 * sizeof(float)));      /* the same 49 values whoever writes them */
    LBAD_CUDA_TRY(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    *out = guard.release();
    return LBAD_OK;
}

extern "C" void lbadcu_resampler_destroy(lbadcu_resampler* r) {
    if (!r) return;
    DeviceScope _device_scope(r->device);
    if (r->stream) { cudaStreamSynchronize(r->stream); cudaStreamDestroy(r->stream); }
    cudaFree(r->d_g); cudaFree(r->d_hc);
    delete r;
}

extern "C" uint64_t lbadcu_resampler_launches(const lbadcu_resampler* r) { return r ? r->launches : 0; }

extern "C" int lbadcu_resample_device(lbadcu_resampler* r, const float* d_in, uint32_t n_clips, uint64_t in_len, uint64_t in_stride,
                                      float* d_out, uint64_t out_len, uint64_t out_stride, void* stream) {
    if (!r || !d_in || !d_out || n_clips == 0) return LBAD_ERR_ARG;
    if (out_len == 0) return LBAD_OK;
    LBAD_ON_DEVICE(r->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : r->stream;
    ResampleParams P = r->P;
    P.in_len = in_len; P.in_stride = in_stride; P.out_len = out_len; P.out_stride = out_stride;
    P.tiles_per_clip = (uint32_t)((out_len + P.tile - 1) / P.tile);
    const uint64_t total = (uint64_t)P.tiles_per_clip * n_clips;
    if (out_len >= (1ull << 31) || total >= (1ull << 31)) return LBAD_ERR_ARG;
    const size_t smem = r->smem_base * sizeof(float);
    const bool d4 = P.D == 4 && P.T1 == 49;
    auto kern = d4 ? (P.tile == 2 * RS_THREADS ? resample_kernel<true, 42, 2> : P.T2 == 42 ? resample_kernel<true, 42, 1> : resample_kernel<true, 0, 1>) : resample_kernel<false, 0, 1>;
    const int vec_ok = ((uintptr_t)d_in % 16 == 0) && (in_stride % 4 == 0);
    LBAD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    LBAD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RS_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const uint64_t cap = (uint64_t)r->sm_count * per_sm;
    const uint32_t grid = (uint32_t)(total < cap ? total : cap);
    kern<<<grid, RS_THREADS, smem, s>>>(d_in, d_out, r->d_g, r->d_hc, P, total, vec_ok);
    r->launches++;
    LBAD_CUDA_TRY(cudaGetLastError());
    return LBAD_OK;
}

extern "C" int lbadcu_resample_host(lbadcu_resampler* r, const float* h_in, uint64_t n_in, float* h_out, uint64_t n_out) {
    if (!r || !h_in || !h_out) return LBAD_ERR_ARG;
    if (n_out == 0) return LBAD_OK;
    LBAD_ON_DEVICE(r->device);
    DevBuf<float> d_in, d_out;
    LBAD_CUDA_TRY(d_in.alloc(n_in)); LBAD_CUDA_TRY(d_out.alloc(n_out));
    LBAD_CUDA_TRY(cudaMemcpyAsync(d_in, h_in, n_in * sizeof(float), cudaMemcpyHostToDevice, r->stream));
    int e = lbadcu_resample_device(r, d_in, 1, n_in, n_in, d_out, n_out, n_out, r->stream);
    if (e != LBAD_OK) return e;
    LBAD_CUDA_TRY(cudaMemcpyAsync(h_out, d_out, n_out * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
    LBAD_CUDA_TRY(cudaStreamSynchronize(r->stream));
    return LBAD_OK;
}

extern "C" int lbadcu_process_recorded_host(lbadcu_plan* p, lbadcu_resampler* r, const float* h_in, uint64_t n_in, uint64_t out_len, uint32_t* h_words, size_t n_words) {
    if (!p || !r || !h_in || !h_words) return LBAD_ERR_ARG;
    LBAD_ON_DEVICE(r->device);
    cudaStream_t s = (cudaStream_t)lbadcu_plan_stream(p);
    DevBuf<float> d_in, d_mid; DevBuf<uint32_t> d_words;
    LBAD_CUDA_TRY(d_in.alloc(n_in)); LBAD_CUDA_TRY(d_mid.alloc(out_len + 4)); LBAD_CUDA_TRY(d_words.alloc(n_words));
    LBAD_CUDA_TRY(cudaMemcpyAsync(d_in, h_in, n_in * sizeof(float), cudaMemcpyHostToDevice, s));
    int e = lbadcu_resample_device(r, d_in, 1, n_in, n_in, d_mid, out_len, out_len, s);
    if (e == LBAD_OK) e = lbadcu_extract_device(p, d_mid, 1, out_len, out_len, d_words, nullptr, nullptr, 0, s);
    if (e != LBAD_OK) { cudaStreamSynchronize(s); return e; }
    LBAD_CUDA_TRY(cudaMemcpyAsync(h_words, d_words, n_words * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    LBAD_CUDA_TRY(cudaStreamSynchronize(s));
    return LBAD_OK;
}

extern "C" int lbadcu_device_alloc_floats(uint64_t n, float** out) {
    *out = nullptr;
    LBAD_CUDA_TRY(cudaMalloc(out, (n ? n : 1) * sizeof(float)));
    return LBAD_OK;
}
extern "C" void lbadcu_device_free(void* p) { if (p) cudaFree(p); }
