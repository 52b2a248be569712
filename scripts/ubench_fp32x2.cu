/*
 * ubench_fp32x2.cu — issue-rate microbenchmarks of the FP32 instructions the FFT kernel is made of (sm_100a):
 * scalar FADD / FMUL / FFMA against the packed FADD2 / FMUL2 / FFMA2 forms, alone and mixed, at the kernel's own
 * occupancy (16 warps per SM).  Prints warp instructions per clock per SM sub-partition and the flop rate.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_bin/ubench_fp32x2 scripts/ubench_fp32x2.cu
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int NACC = 16;
constexpr int UNROLL_ITERS = 8;

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float2* out, int iters, float2 a, float2 b, float2 c) {
    float2 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL_ITERS; u++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) {
                if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }            /* 2 FFMA */
                if (MODE == 1) { acc[i] = __ffma2_rn(acc[i], a, b); }                                                   /* 1 FFMA2 */
                if (MODE == 2) { acc[i].x = acc[i].x + a.x; acc[i].y = acc[i].y + a.y; }                                /* 2 FADD */
                if (MODE == 3) { acc[i] = __fadd2_rn(acc[i], a); }                                                      /* 1 FADD2 */
                if (MODE == 4) { acc[i].x = acc[i].x * a.x; acc[i].y = acc[i].y * a.y; }                                /* 2 FMUL */
                if (MODE == 5) { acc[i] = __fmul2_rn(acc[i], a); }                                                      /* 1 FMUL2 */
                if (MODE == 6) { acc[i] = __ffma2_rn(acc[i], acc[(i + 5) % NACC], acc[(i + 11) % NACC]); }              /* FFMA2, three varying operands */
                if (MODE == 7) { acc[i].x = fmaf(acc[i].x, acc[(i + 5) % NACC].y, acc[(i + 11) % NACC].x); acc[i].y = fmaf(acc[i].y, acc[(i + 3) % NACC].x, acc[(i + 7) % NACC].y); }
                if (MODE == 8) { if (i & 1) acc[i] = __ffma2_rn(acc[i], a, b); else { acc[i].x = acc[i].x + a.x; acc[i].y = fmaf(acc[i].y, a.y, b.y); } }   /* 1 FFMA2 : 1 FADD : 1 FFMA */
                if (MODE == 9) { acc[i] = __ffma2_rn(acc[i], a, b); acc[i].x = fmaxf(acc[i].x, c.x); }                  /* FFMA2 + FMNMX (alu pipe) */
                if (MODE == 10) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaxf(acc[i].y, c.y); }               /* FFMA + FMNMX */
                if (MODE == 11) { acc[i] = __fadd2_rn(acc[i], acc[(i + 5) % NACC]); }                                   /* FADD2, two varying operands */
                if (MODE == 12) { acc[i].x = fmaf(acc[i].x, 1.0009765625f, 0.5f); acc[i].y = fmaf(acc[i].y, 0.9990234375f, -0.5f); }   /* FFMA immediate forms */
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NACC; i++) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Case { const char* name; double inst_per_slot; double flop_per_slot; };

template <int MODE>
static void run(const Case& cs, int sms, float2* d_out, double clk_ghz) {
    const int iters = 2000;
    const int grid = sms * 2;
    float2 a = make_float2(1.0000001f, 0.9999999f), b = make_float2(1e-7f, -1e-7f), c = make_float2(-1e30f, -1e30f);
    k<MODE><<<grid, 256>>>(d_out, 10, a, b, c);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k<MODE><<<grid, 256>>>(d_out, iters, a, b, c);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double slots = (double)iters * UNROLL_ITERS * NACC;            /* per thread */
    const double warp_inst = slots * cs.inst_per_slot * (256 / 32) * grid;
    const double clocks = ms * 1e-3 * clk_ghz * 1e9;
    printf("%-52s %8.3f ms  %6.3f warp-inst/clk/SMSP  %7.2f Tflop/s\n", cs.name, ms, warp_inst / clocks / (sms * 4),
           slots * cs.flop_per_slot * 256.0 * grid / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double ghz = clk_khz * 1e-6;
    printf("%s, %d SMs, clock %.3f GHz (nominal; rates below assume it)\n", p.name, p.multiProcessorCount, ghz);
    float2* d; CK(cudaMalloc(&d, sizeof(float2) * 256 * p.multiProcessorCount * 2));
    const int sms = p.multiProcessorCount;
    run<0>({"2 FFMA (reg, reg, reg)", 2, 4}, sms, d, ghz);
    run<12>({"2 FFMA (immediate forms)", 2, 4}, sms, d, ghz);
    run<1>({"1 FFMA2 (acc, const pair, const pair)", 1, 4}, sms, d, ghz);
    run<6>({"1 FFMA2 (three varying pairs)", 1, 4}, sms, d, ghz);
    run<7>({"2 FFMA (three varying registers)", 2, 4}, sms, d, ghz);
    run<2>({"2 FADD", 2, 2}, sms, d, ghz);
    run<3>({"1 FADD2", 1, 2}, sms, d, ghz);
    run<11>({"1 FADD2 (two varying pairs)", 1, 2}, sms, d, ghz);
    run<4>({"2 FMUL", 2, 2}, sms, d, ghz);
    run<5>({"1 FMUL2", 1, 2}, sms, d, ghz);
    run<8>({"mix: 1 FFMA2 : 1 FADD : 1 FFMA per 2 slots", 1.5, 3.5}, sms, d, ghz);
    run<9>({"1 FFMA2 + 1 FMNMX", 2, 4}, sms, d, ghz);
    run<10>({"1 FFMA + 1 FMNMX", 2, 2}, sms, d, ghz);
    return 0;
}
