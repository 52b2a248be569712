/*
 * lane_emulator.cpp — TEST SUPPORT (host only, g++): runs the fused kernel's per-window warp procedure
 * (lbad_extract.cu: extract_fused_kernel, "32 windows per warp" loop) lane by lane on the CPU, using the very same
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
    constexpr int SCR_LD = 34;
    static float tw1[1024][2], tw2[1024][2];
    static bool init = false;
    if (!init) {
        for (int k1 = 0; k1 < 32; k1++) for (int l = 0; l < 32; l++) { double a = 2.0 * M_PI * (double)(l * k1) / 1024.0; tw1[k1 * 32 + l][0] = (float)cos(a); tw1[k1 * 32 + l][1] = (float)-sin(a); }
        for (int k2 = 0; k2 < 32; k2++) for (int l = 0; l < 32; l++) { double a = 2.0 * M_PI * (double)(l + 32 * k2) / 2048.0; tw2[k2 * 32 + l][0] = (float)cos(a); tw2[k2 * 32 + l][1] = (float)sin(a); }
        init = true;
    }
    static float re[32][32], im[32][32];            /* [lane][register] */
    std::vector<float> scr(32 * SCR_LD * 2), vbuf(1024, 0.0f);
    for (int lane = 0; lane < 32; lane++) {
        for (int n1 = 0; n1 < 32; n1++) { re[lane][n1] = win[2 * (32 * n1 + lane)]; im[lane][n1] = win[2 * (32 * n1 + lane) + 1]; }
        fft32(re[lane], im[lane]);
        for (int p = 0; p < 32; p++) {
            const int k1 = bitrev5(p);
            const float wx = tw1[k1 * 32 + lane][0], wy = tw1[k1 * 32 + lane][1];
            scr[2 * (k1 * SCR_LD + lane)] = re[lane][p] * wx - im[lane][p] * wy;
            scr[2 * (k1 * SCR_LD + lane) + 1] = re[lane][p] * wy + im[lane][p] * wx;
        }
    }
    for (int lane = 0; lane < 32; lane++) {
        for (int q = 0; q < 16; q++) {
            const float* t = &scr[2 * (lane * SCR_LD + 2 * q)];
            re[lane][2 * q] = t[0]; im[lane][2 * q] = t[1]; re[lane][2 * q + 1] = t[2]; im[lane][2 * q + 1] = t[3];
        }
        fft32(re[lane], im[lane]);
    }
    const int k2lo = (int)(kmin >> 5), k2hi = (int)((kmax - 1) >> 5);
    for (int lane = 0; lane < 32; lane++) {
        const int src_lane = (32 - lane) & 31;
        for (int k2 = 0; k2 < 32; k2++) {
            if (k2 >= k2lo && k2 <= k2hi) {
                const int p = bitrev5(k2), pp = bitrev5(31 - k2), p0 = bitrev5((32 - k2) & 31);
                float pr = re[src_lane][pp], pi = im[src_lane][pp];      /* __shfl_sync */
                if (lane == 0) { pr = re[lane][p0]; pi = im[lane][p0]; }
                float xr, xi;
                real_split_2x(re[lane][p], im[lane][p], pr, pi, tw2[k2 * 32 + lane][0], tw2[k2 * 32 + lane][1], xr, xi);
                if (k2 == 0 && lane == 0) { xr = 2.0f * (re[lane][p] + im[lane][p]); xi = 2.0f * (re[lane][p] - im[lane][p]); }
                if (out_spec) { out_spec[2 * (k2 * 32 + lane)] = xr; out_spec[2 * (k2 * 32 + lane) + 1] = xi; }
                vbuf[k2 * 32 + lane] = bin_energy(xr, xi, inv_pos_scale);
            }
        }
    }
    for (int lane = 0; lane < 32; lane++) {
        float pacc = 0.0f;
        for (uint32_t k = klow[lane]; k < khigh[lane]; k++) pacc = pacc + vbuf[k];
        out_bands[lane] = pacc / divisor[lane];
    }
}

/* bare 32-point DFT in natural order, for a direct unit test of fft32 + bitrev5 */
extern "C" void lbad_emulate_fft32(const float* in_re, const float* in_im, float* out_re, float* out_im) {
    float re[32], im[32];
    memcpy(re, in_re, sizeof re); memcpy(im, in_im, sizeof im);
    fft32(re, im);
    for (int p = 0; p < 32; p++) { out_re[bitrev5(p)] = re[p]; out_im[bitrev5(p)] = im[p]; }
}
