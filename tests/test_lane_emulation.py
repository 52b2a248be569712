"""The fused kernel's per-window warp procedure, emulated lane by lane on the CPU with the same lbad_math.cuh code
and index expressions (tests/lane_emulator.cpp), against the oracle: checks the FFT/transposition/split algebra."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from oracle.oracle import Cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lbaudiodetective_b200", "csrc")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "liblane.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-include", "cstdint",
                    os.path.join(ROOT, "tests", "lane_emulator.cpp"), "-I", CSRC, "-o", so], check=True)
    return C.CDLL(so)


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_fft32_natural_order(emu):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(32) + 1j * rng.standard_normal(32))
    ir, ii = x.real.astype(np.float32), x.imag.astype(np.float32); orr, oi = np.zeros(32, np.float32), np.zeros(32, np.float32)
    emu.lbad_emulate_fft32(vp(ir), vp(ii), vp(orr), vp(oi))
    ref = np.fft.fft(ir.astype(np.float64) + 1j * ii.astype(np.float64))
    assert np.abs(orr + 1j * oi - ref).max() < 5e-6


def test_window_pipeline_matches_oracle(emu, port):
    cfg = Cfg.default(); idx, lo, hi = port.band_table(cfg); div = (idx[1:] - idx[:-1]).astype(np.float32)
    pcm = port.synth_clip(3, 55120)
    for w in range(0, 700, 53):
        win = np.ascontiguousarray(pcm[64 * w:64 * w + 2048]); out = np.zeros(32, np.float32); spec = np.zeros(2048, np.float32)
        emu.lbad_emulate_window(vp(win), vp(lo), vp(hi), vp(div), C.c_float(1 / 512.0), C.c_uint32(int(lo.min())), C.c_uint32(int(hi.max())), vp(out), vp(spec))
        ref_spec = port.fft2x(win).reshape(-1, 2); k = np.arange(lo.min(), hi.max())
        assert np.abs(spec.reshape(-1, 2)[k] - ref_spec[k]).max() <= 2e-6 * np.abs(ref_spec).max()
        ref_bands = port.band_energies(cfg, pcm[64 * w:], 1)[0]
        assert (np.abs(out - ref_bands) / np.abs(ref_bands)).max() < 1e-4


@pytest.mark.parametrize("window", [2048, 1024, 512, 256])
def test_multi_window_warp_matches_oracle(emu, port, window):
    """bands_fused_kernel<R>: 32/R windows of 64 R samples side by side in one warp — same index algebra, checked for every R."""
    R = window // 64; S = 32 // R
    cfg = Cfg.default(window=window); idx, lo, hi = port.band_table(cfg); div = (idx[1:] - idx[:-1]).astype(np.float32)
    pcm = port.synth_clip(4, 30000)
    for start in (0, 64 * 7, 64 * 100):
        seg = np.ascontiguousarray(pcm[start:start + 64 * (S - 1) + window])
        out = np.zeros((S, 32), np.float32); spec = np.zeros((S, window), np.float32)
        emu.lbad_emulate_windows(C.c_int(R), vp(seg), C.c_int(64), vp(lo), vp(hi), vp(div), C.c_float(4.0 / window), C.c_uint32(int(lo.min())), C.c_uint32(int(hi.max())), vp(out), vp(spec))
        for s in range(S):
            win = seg[64 * s: 64 * s + window]
            ref_spec = port.fft2x(win).reshape(-1, 2); k = np.arange(lo.min(), hi.max())
            assert np.abs(spec[s].reshape(-1, 2)[k] - ref_spec[k]).max() <= 2e-6 * np.abs(ref_spec).max(), (window, s)
            ref_bands = port.band_energies(cfg, seg[64 * s:], 1)[0]
            floor = 1e-9 * np.abs(ref_bands).max()
            assert (np.abs(out[s] - ref_bands) <= 3e-4 * np.abs(ref_bands) + floor).all(), (window, s)


def test_dc_and_nyquist_packing(emu, port):
    """bin 0 carries 2X[0] in re and 2X[N/2] in im (vDSP packing)."""
    rng = np.random.default_rng(5); win = rng.standard_normal(2048).astype(np.float32)
    lo = np.zeros(32, np.uint32); hi = np.full(32, 40, np.uint32); div = np.ones(32, np.float32); out = np.zeros(32, np.float32); spec = np.zeros(2048, np.float32)
    emu.lbad_emulate_window(vp(win), vp(lo), vp(hi), vp(div), C.c_float(1 / 512.0), C.c_uint32(0), C.c_uint32(40), vp(out), vp(spec))
    ref = port.fft2x(win)
    assert np.abs(spec[:80] - ref[:80]).max() <= 2e-6 * np.abs(ref).max()


def test_bin_energy_matches_literal_formula(emu):
    """fma(max(x,0), 1/s-1, x) == x/s for x > 0 and x otherwise, for the power-of-two scales the reference produces."""
    emu.lbad_emulate_bin_energy.restype = C.c_float
    emu.lbad_emulate_bin_energy.argtypes = [C.c_float, C.c_float, C.c_float]
    rng = np.random.default_rng(9)
    vals = np.concatenate([rng.standard_normal(2000).astype(np.float32) * np.float32(10.0) ** rng.integers(-30, 30, 2000).astype(np.float32),
                           np.array([0.0, -0.0, 1e-45, -1e-45, 3e38, -3e38, np.inf, -np.inf, np.nan], np.float32)])
    for scale in (64.0, 128.0, 256.0, 512.0):
        for re, im in zip(vals, vals[::-1]):
            r = np.float32(re / np.float32(scale)) if re > 0 else np.float32(re)
            i = np.float32(im / np.float32(scale)) if im > 0 else np.float32(im)
            with np.errstate(all="ignore"):
                v = np.float32(np.float32(r * r) + np.float32(i * i))
                want = v if np.isfinite(v) else np.float32(0)
                got = np.float32(emu.lbad_emulate_bin_energy(float(re), float(im), scale))
            # the kernel fuses re*re + im*im into one FMA: allow one rounding of difference on the sum
            assert got == want or abs(float(got) - float(want)) <= 1.2e-7 * abs(float(want)), (re, im, scale, got, want)
