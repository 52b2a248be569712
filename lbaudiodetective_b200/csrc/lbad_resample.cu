/*
 * lbad_resample.cu — recording-rate -> processing-rate conversion on the device (sm_100a), the step ExtAudioFile's client
 * format performs in the reference (LBAudioDetective.m:229, m:275; AudioToolbox itself is out of scope).  The conversion is
 * the two-stage band-limited decimator defined in include/LBAudioDetectiveResample.h; the tables are designed on the host
 * (lbad_resample_design.c) and this file evaluates them: one CTA per tile of 256 output samples of one clip, the input span of
 * the tile staged once in shared memory (every recorded sample is read from HBM once per tile it touches; neighbouring tiles
 * overlap by the filter length, served by L2), stage 1 into a second shared tile, stage 2 from there.  Sums are float32 FMA
 * chains in increasing tap order, so the result equals oracle/lbad_oracle.c bit for bit.
 */
#include "lbad_common.cuh"
#include <math.h>

namespace lbad {

constexpr int RS_THREADS = 256;              /* = output samples per tile */
constexpr int RS_PHASES = LBAD_RS_PHASES;

struct ResampleParams {
    uint32_t D, H1, T1, H2, T2;
    double rho2;
    uint64_t in_len, in_stride, out_len, out_stride;
    uint32_t tiles_per_clip, ny_max, nx_max;
    uint32_t nip, xs_floats;         /* D = 4: floats per polyphase plane of the staged input; floats reserved for the staged input */
};

/* The stage-1 taps for D = 4 depend on nothing but D (sinc((t - 24) / 4) / 4 under a Kaiser window), so one constant copy serves every
 * resampler of the process; as constant-bank operands of the FMAs they cost neither registers nor loads. */
__constant__ float c_g_d4[49];

/* D4: integer decimation by 4 with the 49-tap stage-1 filter — the 44.1 kHz case.  The tile's input starts on a multiple of 4
 * samples, so thread j reads its 49 inputs as twelve conflict-free 128-bit shared loads plus one scalar.  T2C: the number of
 * stage-2 taps when known at compile time (42 for 44.1 kHz -> 5512 Hz), 0 = run-time.  vec_ok: the clips start on 16-byte
 * boundaries, so interior tiles are staged with 128-bit loads. */
template <bool D4, int T2C>
__global__ void __launch_bounds__(RS_THREADS)
resample_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ g, const float* __restrict__ hc, const ResampleParams P,
                const uint64_t total_tiles, const int vec_ok) {
    extern __shared__ __align__(16) float rs_smem[];
    float* xs = rs_smem;                                       /* [nx_max], 16-byte aligned */
    float* y1s = xs + P.xs_floats;                             /* [ny_max + 4] */
    float* gs = y1s + ((P.ny_max + 4 + 3) & ~3u);              /* [T1] */
    float* hcs = gs + ((P.T1 + 3) & ~3u);                      /* [PHASES + 1][T2P], rows padded to T2P = a multiple of 4 floats (zeros) */
    const uint32_t T2 = T2C ? (uint32_t)T2C : P.T2;
    const uint32_t T2P = (T2 + 3) & ~3u;
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < P.T1; i += RS_THREADS) gs[i] = g[i];
    for (uint32_t i = tid; i < (RS_PHASES + 1) * T2P; i += RS_THREADS) { const uint32_t r = i / T2P, c = i % T2P; hcs[i] = c < T2 ? __ldg(hc + r * T2 + c) : 0.0f; }
    for (uint32_t tile = blockIdx.x; tile < (uint32_t)total_tiles; tile += gridDim.x) {      /* the host keeps total_tiles and out_len below 2^31 */
        const uint32_t clip = tile / P.tiles_per_clip;
        const uint32_t m0 = (tile % P.tiles_per_clip) * RS_THREADS;
        const uint32_t m1 = (m0 + RS_THREADS < (uint32_t)P.out_len) ? m0 + RS_THREADS : (uint32_t)P.out_len;
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
                 * of 4): a thread then produces FOUR consecutive stage-1 outputs from four 128-bit loads per plane, 16 loads for 196 FMAs,
                 * instead of 13 loads per output — stage 1 was bound by shared-memory wavefronts */
                const uint32_t ni = (nx + 3) / 4, nip = P.nip;                              /* plane length used / allocated (multiple of 4 floats, >= ny_max + 16) */
                if (vec_ok && x_lo >= 0 && (uint64_t)x_lo + 4ull * ni <= P.in_len) {
                    const float4* s4 = reinterpret_cast<const float4*>(src + x_lo);
                    for (uint32_t i = tid; i < ni; i += RS_THREADS) { const float4 v = __ldg(s4 + i); xs[i] = v.x; xs[nip + i] = v.y; xs[2 * nip + i] = v.z; xs[3 * nip + i] = v.w; }
                } else {
                    for (uint32_t i = tid; i < 4 * ni; i += RS_THREADS) { const int64_t k = x_lo + i; xs[(i & 3) * nip + (i >> 2)] = (k >= 0 && (uint64_t)k < P.in_len) ? __ldg(src + k) : 0.0f; }
                }
                __syncthreads();
                for (uint32_t j0 = 4 * tid; j0 < ny; j0 += 4 * RS_THREADS) {             /* stage 1: outputs j0 .. j0 + 3 */
                    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
                    for (int pl = 0; pl < 4; pl++) {
                        const float4* x4 = reinterpret_cast<const float4*>(xs + pl * nip) + (j0 >> 2);
                        const float4 q0 = x4[0], q1 = x4[1], q2 = x4[2], q3 = x4[3];
                        const float v[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
                        for (int m = 0; m < (pl == 0 ? 13 : 12); m++) {                   /* tap t = 4 m + pl multiplies plane sample (output + m) */
                            const float c = c_g_d4[4 * m + pl];
                            a0 = fmaf(c, v[m], a0); a1 = fmaf(c, v[m + 1], a1); a2 = fmaf(c, v[m + 2], a2); a3 = fmaf(c, v[m + 3], a3);
                        }
                    }
                    *reinterpret_cast<float4*>(y1s + j0) = make_float4(a0, a1, a2, a3);    /* y1s has room for the up to three outputs past ny */
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
        const uint32_t m = m0 + tid;
        if (m < m1) {                                           /* stage 2: position in the stage-1 sequence, coarse phase, interpolation weight */
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
    LBAD_CUDA_TRY(cudaGetDevice(&r->device));
    cudaDeviceProp prop; LBAD_CUDA_TRY(cudaGetDeviceProperties(&prop, r->device));
    r->sm_count = prop.multiProcessorCount;
    ResampleParams& P = r->P;
    P.D = d->D; P.H1 = d->H1; P.T1 = d->D > 1 ? d->T1 : 0; P.H2 = d->H2; P.T2 = d->T2; P.rho2 = d->rho2;
    P.ny_max = (uint32_t)floor(RS_THREADS * d->rho2) + 2 * d->H2 + 2;
    P.nx_max = d->D > 1 ? d->D * (P.ny_max - 1) + d->T1 : 0;
    P.nip = (P.nx_max / 4 + 8) & ~3u;
    P.xs_floats = (d->D == 4 && d->T1 == 49) ? 4 * P.nip : ((P.nx_max + 3) & ~3u);
    const size_t nhc = (size_t)(RS_PHASES + 1) * d->T2;
    /* shared layout: input tile, stage-1 tile (+4: the 64-bit sample loads of stage 2 may read up to three floats past it), taps, padded coarse-phase rows */
    r->smem_base = (size_t)P.xs_floats + (((size_t)P.ny_max + 4 + 3) & ~3u) + (((size_t)P.T1 + 3) & ~3u) + (size_t)(RS_PHASES + 1) * ((d->T2 + 3) & ~3u);
    if (r->smem_base * sizeof(float) > prop.sharedMemPerBlockOptin) { delete r; set_error("resampler: the filters for this rate pair do not fit in shared memory"); return LBAD_ERR_ARG; }
    LBAD_CUDA_TRY(cudaMalloc(&r->d_hc, nhc * sizeof(float)));
    LBAD_CUDA_TRY(cudaMemcpy(r->d_hc, d->hc, nhc * sizeof(float), cudaMemcpyHostToDevice));
    LBAD_CUDA_TRY(cudaMalloc(&r->d_g, (P.T1 ? P.T1 : 1) * sizeof(float)));
    if (P.T1) LBAD_CUDA_TRY(cudaMemcpy(r->d_g, d->g, P.T1 * sizeof(float), cudaMemcpyHostToDevice));
    if (P.D == 4 && P.T1 == 49) LBAD_CUDA_TRY(cudaMemcpyToSymbol(c_g_d4, d->g, 49 * sizeof(float)));      /* the same 49 values whoever writes them */
    LBAD_CUDA_TRY(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    *out = r;
    return LBAD_OK;
}

extern "C" void lbadcu_resampler_destroy(lbadcu_resampler* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->stream) { cudaStreamSynchronize(r->stream); cudaStreamDestroy(r->stream); }
    cudaFree(r->d_g); cudaFree(r->d_hc);
    delete r;
}

extern "C" uint64_t lbadcu_resampler_launches(const lbadcu_resampler* r) { return r ? r->launches : 0; }

extern "C" int lbadcu_resample_device(lbadcu_resampler* r, const float* d_in, uint32_t n_clips, uint64_t in_len, uint64_t in_stride,
                                      float* d_out, uint64_t out_len, uint64_t out_stride, void* stream) {
    if (!r || !d_in || !d_out || n_clips == 0) return LBAD_ERR_ARG;
    if (out_len == 0) return LBAD_OK;
    LBAD_CUDA_TRY(cudaSetDevice(r->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : r->stream;
    ResampleParams P = r->P;
    P.in_len = in_len; P.in_stride = in_stride; P.out_len = out_len; P.out_stride = out_stride;
    P.tiles_per_clip = (uint32_t)((out_len + RS_THREADS - 1) / RS_THREADS);
    const uint64_t total = (uint64_t)P.tiles_per_clip * n_clips;
    if (out_len >= (1ull << 31) || total >= (1ull << 31)) return LBAD_ERR_ARG;
    const size_t smem = r->smem_base * sizeof(float);
    const bool d4 = P.D == 4 && P.T1 == 49;
    auto kern = d4 ? (P.T2 == 42 ? resample_kernel<true, 42> : resample_kernel<true, 0>) : resample_kernel<false, 0>;
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
    LBAD_CUDA_TRY(cudaSetDevice(r->device));
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
    LBAD_CUDA_TRY(cudaSetDevice(r->device));
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
