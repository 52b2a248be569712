"""The C restatement (oracle/lbad_oracle.c) against the committed outputs of the compiled reference (tests/golden/)."""
import numpy as np
from oracle.oracle import Cfg


def test_band_tables(port, kat):
    for n, t in kat["band_tables"].items():
        idx, lo, hi = port.band_table(Cfg.default(window=int(n)))
        assert idx.tolist() == t["indices"] and lo.tolist() == t["klow"] and hi.tolist() == t["khigh"]
    # SURVEY.md §2.2 exact tables at defaults
    idx, lo, hi = port.band_table(Cfg.default())
    assert (lo[0], hi[-1], idx[0], idx[-1]) == (86, 759, 118, 1023)
    assert int((hi - lo).sum()) == 673


def test_haar_known_answer(port, kat):
    out = port.haar(np.array(kat["haar_3x4_in"], np.float32))
    assert np.array_equal(out, np.array(kat["haar_3x4_out"], np.float32))
    assert abs(out[0, 0] - 969.385559) < 1e-3 and abs(out[2, 3] + 291.693420) < 1e-3


def test_compare_toy(port, kat):
    t = kat["compare_toy"]; a = np.array(t["a"]); b = np.array(t["b"])
    assert port.compare_sub(a, b, 8, 8) == t["ab8"] == np.float32(2 / 3)
    assert port.compare_sub(b, a, 8, 8) == t["ba8"] == 0.5
    assert port.compare_sub(a, b, 8, 4) == t["ab4"] == 0.5


def test_config1_stages_bit_identical(port, config1):
    cfg = Cfg.default()
    for c in range(2):
        bits, img, haar = port.process(cfg, config1["pcm"][c], stages=True)
        assert np.array_equal(img, config1["images"][c])
        assert np.array_equal(haar, config1["haar"][c])
        assert np.array_equal(bits, config1["bits"][c])
    assert bits.shape == (6, 200)


def test_config1_scores(port, config1):
    cfg = Cfg.default(); p = config1["pcm"]
    assert np.float32(port.compare_pcm(cfg, p[0], p[1], 0)) == config1["score_01"]
    assert np.float32(port.compare_pcm(cfg, p[1], p[0], 0)) == config1["score_10"]
    assert np.float32(port.compare_pcm(cfg, p[0], p[0], 0)) == config1["score_00"] == 1.0
    assert np.float32(port.compare_pcm(cfg, p[0], p[1], 100)) == config1["score_01_r100"]


def test_compare_cases(port, compare_cases):
    for c in compare_cases:
        assert np.float32(port.compare_fp(c["fp1"], c["fp2"], int(c["range"]))) == c["score"], (int(c["c1"]), int(c["c2"]), int(c["range"]))


def test_sweep(port, sweep):
    for n in (512, 1024, 2048):
        for L in (100, 200, 400):
            cfg = Cfg.default(window=n, sublen=L)
            assert np.array_equal(port.process(cfg, sweep["base"]), sweep["clip_%d_%d" % (n, L)])
            assert np.array_equal(port.process(cfg, sweep["query"]), sweep["query_%d_%d" % (n, L)])


def test_subfp_counts(port, kat):
    from oracle.oracle import subfp_count
    for n, c in kat["subfp_counts"].items():
        assert subfp_count(Cfg.default(), int(n)) == c
    assert subfp_count(Cfg.default(), 2047) == 0 and subfp_count(Cfg.default(), 10239) == 0 and subfp_count(Cfg.default(), 10240) == 1
