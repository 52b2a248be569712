/*
 * lane_emulator.cpp — TEST SUPPORT (host only, g++): runs the fused kernel's per-window warp procedure
 * (lbad_extract.cu: extract_fused_kernel, "16 windows per warp" loop) lane by lane on the CPU, using the very same
 * lbad_math.cuh functions and the same index expressions, with shared memory and shuffles replaced by arrays.
 * It lets the register/shared-memory index algebra of the kernel be checked against the oracle without a GPU
 * (tests/test_lane_emulation.py).  It is not part of libLBAudioDetectiveCUDA.so and never ships.
 */
#include "lbad_math.cuh"
#include <cmath>
#include <cstring>
#include <vector>

using namespace lbad;

extern "C" void lbad_emulate_window(const float* win, const uint32_t* klow, const uint32_t* khigh, const float* divisor,
                                    float inv_pos_scale, uint32_t kmin, uint32_t kmax, float* out_bands, float* out_spec /* 2048 floats or NULL */) {
    constexpr int SCR_LDF = 36;
    static float tw1[512][4], tw2[512][4];
    static bool init = false;
    if (!init) {                                   /* same tables as lbadcu_plan_create */
        for (int pp = 0; pp < 32; pp += 2) for (int l = 0; l < 32; l++) {
            const double a0 = 2.0 * M_PI * (double)(l * bitrev5(pp)) / 1024.0, a1 = 2.0 * M_PI * (double)(l * bitrev5(pp + 1)) / 1024.0;
            float* t = tw1[(pp >> 1) * 32 + l]; t[0] = (float)cos(a0); t[1] = (float)-sin(a0); t[2] = (float)cos(a1); t[3] = (float)-sin(a1);
        }
        for (int k2 = 0; k2 < 32; k2 += 2) for (int l = 0; l < 32; l++) {
            const double a0 = 2.0 * M_PI * (double)(l + 32 * k2) / 2048.0, a1 = 2.0 * M_PI * (double)(l + 32 * (k2 + 1)) / 2048.0;
            float* t = tw2[(k2 >> 1) * 32 + l]; t[0] = (float)cos(a0); t[1] = (float)sin(a0); t[2] = (float)cos(a1); t[3] = (float)sin(a1);
        }
        init = true;
    }
    static float re[32][32], im[32][32];            /* [lane][register] */
    std::vector<float> scr(32 * SCR_LDF), vbuf(1024, 0.0f);
    const float scale_m1 = inv_pos_scale - 1.0f;
    for (int lane = 0; lane < 32; lane++) {
        for (int n1 = 0; n1 < 32; n1++) { re[lane][n1] = win[2 * (32 * n1 + lane)]; im[lane][n1] = win[2 * (32 * n1 + lane) + 1]; }
        fft32(re[lane], im[lane]);
        for (int p = 0; p < 32; p += 2) {
            const float* w = tw1[(p >> 1) * 32 + lane];
            const float a0 = re[lane][p] * w[0] - im[lane][p] * w[1], b0 = re[lane][p] * w[1] + im[lane][p] * w[0];
            const float a1 = re[lane][p + 1] * w[2] - im[lane][p + 1] * w[3], b1 = re[lane][p + 1] * w[3] + im[lane][p + 1] * w[2];
            re[lane][p] = a0; im[lane][p] = b0; re[lane][p + 1] = a1; im[lane][p + 1] = b1;
        }
    }
    for (int comp = 0; comp < 2; comp++) {          /* one component at a time, as in the kernel */
        float (*x)[32] = comp ? im : re;
        for (int lane = 0; lane < 32; lane++) for (int p = 0; p < 32; p++) scr[bitrev5(p) * SCR_LDF + lane] = x[lane][p];
        for (int lane = 0; lane < 32; lane++) for (int q = 0; q < 8; q++) for (int j = 0; j < 4; j++) x[lane][4 * q + j] = scr[lane * SCR_LDF + 4 * q + j];
    }
    for (int lane = 0; lane < 32; lane++) fft32(re[lane], im[lane]);
    const int k2lo = (int)(kmin >> 5), k2hi = (int)((kmax - 1) >> 5);
    for (int lane = 0; lane < 32; lane++) {
        const int src_lane = (32 - lane) & 31;
        for (int k2 = 0; k2 < 32; k2 += 2) {
            if (k2 + 1 >= k2lo && k2 <= k2hi) {
                const float* w = tw2[(k2 >> 1) * 32 + lane];
                for (int h = 0; h < 2; h++) {
                    const int kk = k2 + h;
                    const int p = bitrev5(kk), pp = bitrev5(31 - kk), p0 = bitrev5((32 - kk) & 31);
                    float pr = re[src_lane][pp], pi = im[src_lane][pp];      /* __shfl_sync */
                    if (lane == 0) { pr = re[lane][p0]; pi = im[lane][p0]; }
                    float xr, xi;
                    real_split_2x(re[lane][p], im[lane][p], pr, pi, h ? w[2] : w[0], h ? w[3] : w[1], xr, xi);
                    if (kk == 0 && lane == 0) { xr = 2.0f * (re[lane][p] + im[lane][p]); xi = 2.0f * (re[lane][p] - im[lane][p]); }
                    if (out_spec) { out_spec[2 * (kk * 32 + lane)] = xr; out_spec[2 * (kk * 32 + lane) + 1] = xi; }
                    vbuf[kk * 32 + lane] = bin_energy(xr, xi, scale_m1);
                }
            }
        }
    }
    for (int lane = 0; lane < 32; lane++) {
        float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
        uint32_t k = klow[lane];
        for (; k + 4 <= khigh[lane]; k += 4) { p0 += vbuf[k]; p1 += vbuf[k + 1]; p2 += vbuf[k + 2]; p3 += vbuf[k + 3]; }
        for (; k < khigh[lane]; k++) p0 += vbuf[k];
        out_bands[lane] = ((p0 + p1) + (p2 + p3)) / divisor[lane];
    }
}

/* bare 32-point DFT in natural order, for a direct unit test of fft32 + bitrev5 */
extern "C" void lbad_emulate_fft32(const float* in_re, const float* in_im, float* out_re, float* out_im) {
    float re[32], im[32];
    memcpy(re, in_re, sizeof re); memcpy(im, in_im, sizeof im);
    fft32(re, im);
    for (int p = 0; p < 32; p++) { out_re[bitrev5(p)] = re[p]; out_im[bitrev5(p)] = im[p]; }
}

/* bin_energy against the reference's literal formulation (LBAudioDetective.m:387-401), for the unit test */
extern "C" float lbad_emulate_bin_energy(float re, float im, float pos_scale) { return bin_energy(re, im, 1.0f / pos_scale - 1.0f); }
