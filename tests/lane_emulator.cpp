/*
 * lane_emulator.cpp — TEST SUPPORT (host only, g++): runs the fused kernel's per-window warp procedure
 * (lbaudiodetective_b200/csrc/lbad_extract.cu: bands_fused_kernel<R>, the per-warp window loop) lane by lane on the CPU, using the
 * very same lbad_math.cuh functions and the same index expressions, with shared memory and shuffles replaced by arrays.
 * It lets the register/shared-memory index algebra of the kernel be checked against the oracle without a GPU
 * (tests/test_lane_emulation.py).  It lives with the tests: it is not part of libLBAudioDetectiveCUDA.so and never ships.
 */
#include "lbad_math.cuh"
#include <cmath>
#include <cstring>
#include <vector>

using namespace lbad;

/* One warp iteration of bands_fused_kernel<R, false>: S = 32/R windows of N = 64 R samples, window s starting at pcm + s*hop.
 * out_bands: [S][32]; out_spec (optional): [S][N] interleaved 2X[k] for the bins the kernel computes. */
template <int R>
static void emulate_windows(const float* pcm, int hop, const uint32_t* klow, const uint32_t* khigh, const float* divisor,
                            float inv_pos_scale, uint32_t kmin, uint32_t kmax, float* out_bands, float* out_spec) {
    constexpr int SCR_LDF = 36, S = 32 / R, M = 32 * R, N = 2 * M;
    static float tw1[512][4], tw2[512][4];
    {                                              /* same tables as lbadcu_plan_create */
        for (int pp = 0; pp < 32; pp += 2) for (int l = 0; l < 32; l++) {
            const double a0 = 2.0 * M_PI * (double)((l % R) * bitrev5(pp)) / (double)M, a1 = 2.0 * M_PI * (double)((l % R) * bitrev5(pp + 1)) / (double)M;
            float* t = tw1[(pp >> 1) * 32 + l]; t[0] = (float)cos(a0); t[1] = (float)-sin(a0); t[2] = (float)cos(a1); t[3] = (float)-sin(a1);
        }
        for (int k2 = 0; k2 < 32; k2 += 2) for (int l = 0; l < 32; l++) {
            const double a0 = 2.0 * M_PI * (double)(l + 32 * k2) / (double)N, a1 = 2.0 * M_PI * (double)(l + 32 * (k2 + 1)) / (double)N;
            if (R == 32) {
                float* t = &tw2[0][0] + 2 * (k2 * 32 + l); t[0] = (float)cos(a0); t[1] = (float)sin(a0);
                t = &tw2[0][0] + 2 * ((k2 + 1) * 32 + l); t[0] = (float)cos(a1); t[1] = (float)sin(a1);
            } else {
                float* t = tw2[(k2 >> 1) * 32 + l]; t[0] = (float)cos(a0); t[1] = (float)sin(a0); t[2] = (float)cos(a1); t[3] = (float)sin(a1);
            }
        }
    }
    static float2 z[32][32];                        /* [lane][register] */
    std::vector<float> scr(32 * SCR_LDF), vbuf(1024, 0.0f);
    const float scale_m1 = inv_pos_scale - 1.0f;
    for (int lane = 0; lane < 32; lane++) {
        const int my_win = lane / R, n2 = lane % R;
        const float* win = pcm + (size_t)my_win * hop;
        if constexpr (R == 32) {
            /* the kernel's carried scheme: P = dft16(odd n1) x twiddle is what a window computes; P of the even n1 is the carried P of the
             * window one hop earlier; z[k1] tw[k1] = Pe + W^k1 Po, z[k1+16] tw[k1+16] = (Pe - W^k1 Po) exp(-2 pi i 16 n2 / M) */
            float2 ev[16], od[16];
            for (int m = 0; m < 16; m++) { ev[m] = make_float2(win[2 * (32 * (2 * m) + lane)], win[2 * (32 * (2 * m) + lane) + 1]); od[m] = make_float2(win[2 * (32 * (2 * m + 1) + lane)], win[2 * (32 * (2 * m + 1) + lane) + 1]); }
            dft16(ev); dft16(od);
            for (int q = 0; q < 16; q++) {
                const double a = 2.0 * M_PI * (double)(lane * bitrevR<16>(q)) / (double)M; const float2 t = make_float2((float)cos(a), (float)-sin(a));
                ev[q] = cmul(ev[q], t); od[q] = cmul(od[q], t);
            }
            dit32_combine(ev, od, z[lane]);
            const double ao = 2.0 * M_PI * (double)(16 * lane) / (double)M; const float omx = (float)cos(ao), omy = (float)-sin(ao);
            for (int q = 0; q < 16; q++) { float2& v = z[lane][2 * q + 1]; v = make_float2(v.x * omx - v.y * omy, v.x * omy + v.y * omx); }
        } else {
            for (int n1 = 0; n1 < 32; n1++) z[lane][n1] = make_float2(win[2 * (R * n1 + n2)], win[2 * (R * n1 + n2) + 1]);
            fft32(z[lane]);
            for (int p = 0; p < 32; p += 2) {
                const float* w = tw1[(p >> 1) * 32 + lane];
                float2* zz = z[lane];
                zz[p] = make_float2(zz[p].x * w[0] - zz[p].y * w[1], zz[p].x * w[1] + zz[p].y * w[0]);
                zz[p + 1] = make_float2(zz[p + 1].x * w[2] - zz[p + 1].y * w[3], zz[p + 1].x * w[3] + zz[p + 1].y * w[2]);
            }
        }
    }
    if constexpr (R == 32) {                        /* the carried kernel transposes whole (re, im) pairs: row k1 = bitrev5(p), column lane; lane k1 reads its row */
        constexpr int SCR_LD2 = 34;
        std::vector<float2> scr2(32 * SCR_LD2);
        for (int lane = 0; lane < 32; lane++) for (int p = 0; p < 32; p++) scr2[bitrev5(p) * SCR_LD2 + lane] = z[lane][p];
        for (int lane = 0; lane < 32; lane++) { for (int q = 0; q < 32; q++) z[lane][q] = scr2[lane * SCR_LD2 + q]; fft32_tail<R>(z[lane]); }
    } else {
    static float zx[32][32], zy[32][32];            /* [lane][position]: the transposed components, as the kernel's 128-bit loads deliver them */
    for (int comp = 0; comp < 2; comp++) {          /* one component at a time, as in the kernel */
        for (int lane = 0; lane < 32; lane++) for (int p = 0; p < 32; p++) scr[bitrev5(p) * SCR_LDF + lane] = comp ? z[lane][p].y : z[lane][p].x;
        for (int lane = 0; lane < 32; lane++) for (int q = 0; q < 8; q++) for (int j = 0; j < 4; j++) (comp ? zy[lane][4 * q + j] : zx[lane][4 * q + j]) = scr[lane * SCR_LDF + 4 * q + j];
    }
    for (int lane = 0; lane < 32; lane++) fft32_tail_soa<R>(zx[lane], zy[lane], z[lane]);
    }
    const int k2lo = (int)(kmin >> 5), k2hi = (int)((kmax - 1) >> 5);
    if constexpr (R == 32) {                        /* mirrored rows share one evaluation of the real split (see the kernel) */
        auto row_needed = [&](int r) -> bool { return r >= k2lo && r <= k2hi; };
        auto spec_out = [&](int k, float xr, float xi) { if (out_spec) { out_spec[2 * k] = xr; out_spec[2 * k + 1] = xi; } };
        for (int lane = 0; lane < 32; lane++) {
            const int src_lane = (32 - lane) & 31;
            const float2 wl = make_float2((&tw2[0][0])[2 * lane], (&tw2[0][0])[2 * lane + 1]);      /* the lane's factor: row 0 of the table */
            for (int k2 = 0; k2 < 16; k2++) {
                const bool need_lo = row_needed(k2), need_hi = row_needed(31 - k2) || (k2 > 0 && row_needed(32 - k2));
                if (need_lo || need_hi) {
                    const int p = bitrev5(k2), pp = bitrev5(31 - k2), p0 = bitrev5((32 - k2) % 32);
                    float2 pz = z[src_lane][pp];                             /* __shfl_sync */
                    if (lane == 0) pz = z[lane][p0];
                    const int k = k2 * 32 + lane;
                    float2 lo, hi;
                    real_split_pair_rows(k2, z[lane][p], pz, wl, lo, hi);
                    if (k2 == 0 && lane == 0) { lo.x = 2.0f * (z[lane][p].x + z[lane][p].y); lo.y = 2.0f * (z[lane][p].x - z[lane][p].y); }
                    if (need_lo) { vbuf[k] = bin_energy_raw(lo.x, lo.y, scale_m1); spec_out(k, lo.x, lo.y); }
                    if (need_hi && (k2 > 0 || lane > 0)) { vbuf[1024 - k] = bin_energy_raw_conj(hi.x, hi.y, scale_m1); spec_out(1024 - k, hi.x, -hi.y); }
                }
            }
            if (row_needed(16) && lane == 0) {
                const float xr = 2.0f * z[lane][bitrev5(16)].x, xi = -2.0f * z[lane][bitrev5(16)].y;
                vbuf[512] = bin_energy_raw(xr, xi, scale_m1); spec_out(512, xr, xi);
            }
        }
    } else
    for (int lane = 0; lane < 32; lane++) {
        const int src_lane = (32 - lane) & 31;
        for (int k2 = 0; k2 < R; k2 += 2) {
            if (k2 + 1 >= k2lo && k2 <= k2hi) {
                const float* w = tw2[(k2 >> 1) * 32 + lane];
                for (int h = 0; h < 2; h++) {
                    const int kk = k2 + h;
                    for (int s = 0; s < S; s++) {
                        const int p = s * R + bitrevR<R>(kk), pp = s * R + bitrevR<R>(R - 1 - kk), p0 = s * R + bitrevR<R>((R - kk) % R);
                        float2 pz = z[src_lane][pp];                         /* __shfl_sync */
                        if (lane == 0) pz = z[lane][p0];
                        float xr, xi;
                        real_split_2x(z[lane][p], pz, h ? w[2] : w[0], h ? w[3] : w[1], xr, xi);
                        if (kk == 0 && lane == 0) { xr = 2.0f * (z[lane][p].x + z[lane][p].y); xi = 2.0f * (z[lane][p].x - z[lane][p].y); }
                        if (out_spec) { out_spec[(size_t)s * N + 2 * (kk * 32 + lane)] = xr; out_spec[(size_t)s * N + 2 * (kk * 32 + lane) + 1] = xi; }
                        vbuf[s * M + kk * 32 + lane] = bin_energy_raw(xr, xi, scale_m1);
                    }
                }
            }
        }
    }
    /* band sums as in the kernel: two lanes per band, halves combined by the xor-1 shuffle */
    for (int s = 0; s < S; s++) {
        const float* v = vbuf.data() + s * M;
        auto seg_sum = [&](uint32_t a, uint32_t b) {
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
            for (; a + 4 <= b; a += 4) { s0 += v[a]; s1 += v[a + 1]; s2 += v[a + 2]; s3 += v[a + 3]; }
            if (a + 2 <= b) { s0 += v[a]; s1 += v[a + 1]; a += 2; }
            if (a < b) s2 += v[a];
            return (s0 + s1) + (s2 + s3);
        };
        float sa[32], sb[32];
        for (int lane = 0; lane < 32; lane++) {
            const int b0 = lane >> 1, b1 = 16 + (lane >> 1);
            const uint32_t l0 = klow[b0], h0 = khigh[b0], m0 = l0 + (h0 - l0 + 1) / 2;
            const uint32_t l1 = klow[b1], h1 = khigh[b1], m1 = l1 + (h1 - l1 + 1) / 2;
            sa[lane] = seg_sum((lane & 1) ? m0 : l0, (lane & 1) ? h0 : m0);
            sb[lane] = seg_sum((lane & 1) ? m1 : l1, (lane & 1) ? h1 : m1);
        }
        for (int lane = 0; lane < 32; lane++) {
            const int my_band = (lane >> 1) + ((lane & 1) ? 16 : 0);
            const float a2 = sa[lane] + sa[lane ^ 1], b2 = sb[lane] + sb[lane ^ 1];
            float tot = (lane & 1) ? b2 : a2;
            if (!(tot <= 3.402823466e+38f)) { tot = 0.0f; for (uint32_t k = klow[my_band]; k < khigh[my_band]; k++) if (v[k] <= 3.402823466e+38f) tot += v[k]; }
            out_bands[s * 32 + my_band] = tot / divisor[my_band];
        }
    }
}

extern "C" void lbad_emulate_windows(int R, const float* pcm, int hop, const uint32_t* klow, const uint32_t* khigh, const float* divisor,
                                     float inv_pos_scale, uint32_t kmin, uint32_t kmax, float* out_bands, float* out_spec) {
    if (R == 32) emulate_windows<32>(pcm, hop, klow, khigh, divisor, inv_pos_scale, kmin, kmax, out_bands, out_spec);
    else if (R == 16) emulate_windows<16>(pcm, hop, klow, khigh, divisor, inv_pos_scale, kmin, kmax, out_bands, out_spec);
    else if (R == 8) emulate_windows<8>(pcm, hop, klow, khigh, divisor, inv_pos_scale, kmin, kmax, out_bands, out_spec);
    else if (R == 4) emulate_windows<4>(pcm, hop, klow, khigh, divisor, inv_pos_scale, kmin, kmax, out_bands, out_spec);
}

extern "C" void lbad_emulate_window(const float* win, const uint32_t* klow, const uint32_t* khigh, const float* divisor,
                                    float inv_pos_scale, uint32_t kmin, uint32_t kmax, float* out_bands, float* out_spec /* 2048 floats or NULL */) {
    emulate_windows<32>(win, 0, klow, khigh, divisor, inv_pos_scale, kmin, kmax, out_bands, out_spec);
}

/* bare 32-point DFT in natural order, for a direct unit test of fft32 + bitrev5 */
extern "C" void lbad_emulate_fft32(const float* in_re, const float* in_im, float* out_re, float* out_im) {
    float2 z[32];
    for (int i = 0; i < 32; i++) z[i] = make_float2(in_re[i], in_im[i]);
    fft32(z);
    for (int p = 0; p < 32; p++) { out_re[bitrev5(p)] = z[p].x; out_im[bitrev5(p)] = z[p].y; }
}

/* bin_energy against the reference's literal formulation (LBAudioDetective.m:387-401), for the unit test */
extern "C" float lbad_emulate_bin_energy(float re, float im, float pos_scale) { return bin_energy(re, im, 1.0f / pos_scale - 1.0f); }
