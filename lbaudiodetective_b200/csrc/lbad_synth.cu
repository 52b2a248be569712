/*
 * lbad_synth.cu — bench/test support kernels (not on the reference's path): device-side synthetic PCM with the
 * formula of SURVEY.md §8(d) (host twin: oracle/lbad_oracle.c lbad_synth_clip; device libm, so not bit-identical —
 * parity inputs are always host-generated), random rank-sign codes for search timing, and two microbenchmarks that
 * measure the roofline denominators the fingerprint kernels are actually bound by (FP32 FMA issue, POPC/LOP3 issue).
 */
#include "lbad_common.cuh"

namespace lbad {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline double u01(uint64_t h) { return (double)(h >> 40) * (1.0 / 16777216.0); }

__global__ void synth_kernel(float* __restrict__ out, const uint32_t n_clips, const uint64_t clip_len, const uint64_t clip_stride,
                             const uint64_t first_clip, const uint64_t base_seed, const double sr) {
    const uint32_t clip = blockIdx.y;
    if (clip >= n_clips) return;
    const uint64_t s = splitmix64(base_seed ^ splitmix64(first_clip + clip));
    const double f0 = 250.0 + 350.0 * u01(splitmix64(s + 1));
    const double f1 = 1400.0 + 600.0 * u01(splitmix64(s + 2));
    const double ft = 400.0 + 1400.0 * u01(splitmix64(s + 3));
    const double T = (double)clip_len / sr, kr = (f1 - f0) / (T > 0 ? T : 1.0);
    const uint64_t ns = splitmix64(s + 4);
    float* o = out + (uint64_t)clip * clip_stride;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < clip_len; i += (uint64_t)gridDim.x * blockDim.x) {
        const double t = (double)i / sr;
        double turns = f0 * t + 0.5 * kr * t * t;           /* phase in turns, reduced in double, evaluated in float */
        turns -= floor(turns);
        double tt = ft * t; tt -= floor(tt);
        const float u = (float)(u01(splitmix64(ns + i)) - 0.5);
        float v = 0.5f * sinpif(2.0f * (float)turns) + 0.2f * sinpif(2.0f * (float)tt) + 0.1f * u;
        o[i] = fminf(1.0f, fmaxf(-1.0f, v));
    }
}

/* every one of the first `pairs` ranks carries exactly one sign bit, as extraction produces for non-zero coefficients */
__global__ void random_codes_kernel(uint32_t* __restrict__ words, const uint64_t n_subfps, const uint32_t W, const uint32_t pairs, const uint64_t seed, const uint64_t first_subfp) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_subfps * W; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t sub = i / W; const uint32_t w = (uint32_t)(i % W);
        const uint32_t r = (uint32_t)splitmix64(seed ^ splitmix64(first_subfp * W + i));      /* a function of the GLOBAL word index */
        const uint32_t mask = pairs >= 32u * (w + 1) ? 0xffffffffu : (pairs > 32u * w ? ((1u << (pairs - 32u * w)) - 1u) : 0u);
        words[sub * 2 * W + w] = r & mask;
        words[sub * 2 * W + W + w] = ~r & mask;
    }
}

template <int ILP>
__global__ void fma_bench_kernel(float* out, int iters, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == 123.456f) out[0] = s;
}

template <int ILP, bool POPC>
__global__ void int_bench_kernel(uint32_t* out, int iters, uint32_t a, uint32_t b) {
    uint32_t x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 2654435761u + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (POPC) x[i] = __popc(x[i]) + a;                  /* POPC + IADD */
            else x[i] = (x[i] & a) ^ (x[i] | b) ^ (uint32_t)it;    /* LOP3s */
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}

}  // namespace lbad

using namespace lbad;

extern "C" int lbadcu_synth_device(float* d_out, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride, uint64_t first_clip_id,
                                   uint64_t base_seed, double sample_rate, void* stream) {
    if (!d_out || n_clips == 0 || clip_len == 0) return LBAD_ERR_ARG;
    for (uint32_t c0 = 0; c0 < n_clips; c0 += 32768) {
        const uint32_t nc = n_clips - c0 < 32768 ? n_clips - c0 : 32768;
        dim3 grid((unsigned)((clip_len + 256 * 8 - 1) / (256 * 8)), nc);
        synth_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_out + (uint64_t)c0 * clip_stride, nc, clip_len, clip_stride, first_clip_id + c0, base_seed, sample_rate);
    }
    LBAD_CUDA_TRY(cudaGetLastError());
    return LBAD_OK;
}

extern "C" int lbadcu_random_codes_device(uint32_t* d_words, uint64_t n_subfps, uint32_t W, uint32_t pairs, uint64_t seed, uint64_t first_subfp, void* stream) {
    if (!d_words || n_subfps == 0) return LBAD_ERR_ARG;
    random_codes_kernel<<<2048, 256, 0, (cudaStream_t)stream>>>(d_words, n_subfps, W, pairs, seed, first_subfp);
    LBAD_CUDA_TRY(cudaGetLastError());
    return LBAD_OK;
}

template <class F>
static int time_kernel(F launch, double* ms_out) {
    cudaEvent_t a, b; LBAD_CUDA_TRY(cudaEventCreate(&a)); LBAD_CUDA_TRY(cudaEventCreate(&b));
    launch(); launch();
    LBAD_CUDA_TRY(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        LBAD_CUDA_TRY(cudaEventRecord(a)); launch(); LBAD_CUDA_TRY(cudaEventRecord(b)); LBAD_CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0; LBAD_CUDA_TRY(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *ms_out = best;
    return LBAD_OK;
}

/* fp32_tflops: dependent-FMA chains (2 flop each); popc_gops / lop3_gops: warp-lane operations per second / 1e9 */
extern "C" int lbadcu_microbench(double* fp32_tflops, double* popc_gops, double* lop3_gops) {
    if (lbadcu_device_available() != LBAD_OK) return LBAD_ERR_NODEVICE;
    cudaDeviceProp prop; int dev = 0; LBAD_CUDA_TRY(cudaGetDevice(&dev)); LBAD_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    float* d = nullptr; LBAD_CUDA_TRY(cudaMalloc(&d, 64));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double ms = 0; int e;
    e = time_kernel([&] { fma_bench_kernel<16><<<blocks, threads>>>(d, iters, 1.0000001f, 1e-9f); }, &ms); if (e) return e;
    if (fp32_tflops) *fp32_tflops = 2.0 * 16 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
    e = time_kernel([&] { int_bench_kernel<16, true><<<blocks, threads>>>((uint32_t*)d, iters, 3u, 5u); }, &ms); if (e) return e;
    if (popc_gops) *popc_gops = 16.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e9;
    e = time_kernel([&] { int_bench_kernel<16, false><<<blocks, threads>>>((uint32_t*)d, iters, 0x0f0f0f0fu, 0x33333333u); }, &ms); if (e) return e;
    if (lop3_gops) *lop3_gops = 16.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e9;
    cudaFree(d);
    return LBAD_OK;
}
