"""Recording-rate -> processing-rate conversion (include/LBAudioDetectiveResample.h).  The reference leaves this step to Apple's
ExtAudioFile converter (LBAudioDetective.m:229, m:275), which is closed: PARITY IS UNPINNED against the reference here.  What is
checked: the scalar restatement (oracle/lbad_oracle.c) against analytic tones and scipy's polyphase resampler (CPU), and the CUDA
kernel against the restatement bit for bit (GPU)."""
import numpy as np
import pytest

FS = 44100.0
OUT = 5512.0


def tone(f, n, fs=FS, amp=0.5):
    return (amp * np.sin(2 * np.pi * f * np.arange(n) / fs)).astype(np.float32)


def test_oracle_length_rule(port):
    assert len(port.resample(np.zeros(441000, np.float32), FS)) == 55120          # 10 s -> 10 s
    assert len(port.resample(np.zeros(44100, np.float32), FS)) == 5512
    assert len(port.resample(np.zeros(7, np.float32), FS)) == 0
    assert len(port.resample(np.zeros(5512, np.float32), OUT)) == 5512


def test_oracle_passband_and_stopband(port):
    """Flat over the band the fingerprint reads (231-2043 Hz); whatever could alias into it (above 3469 Hz) is gone."""
    n = 2 * 44100
    for f in (300.0, 1000.0, 2000.0, 2043.0):
        y = port.resample(tone(f, n), FS); m = np.arange(len(y))
        assert np.abs(y[200:-200] - 0.5 * np.sin(2 * np.pi * f * m[200:-200] / OUT)).max() < 2e-3, f
    for f in (3469.0, 3600.0, 5000.0, 8000.0, 12000.0, 20000.0):
        y = port.resample(tone(f, n), FS)
        assert np.abs(y[200:-200]).max() < 0.5 * 10 ** (-80 / 20), f


def test_oracle_against_scipy_polyphase(port):
    """Band-limited noise through scipy.signal.resample_poly (1378/11025, its own Kaiser design) agrees to filter-design accuracy."""
    from scipy.signal import firwin, resample_poly
    rng = np.random.default_rng(0); n = 2 * 44100
    x = np.convolve(rng.standard_normal(n), firwin(1023, 1900.0, fs=FS), mode="same").astype(np.float32)
    y = port.resample(x, FS); ys = resample_poly(x.astype(np.float64), 1378, 11025)[:len(y)]
    assert np.abs(y[300:-300] - ys[300:-300]).max() < 5e-3 * np.abs(ys).max()


def test_oracle_other_rates(port):
    for fs in (48000.0, 22050.0, 16000.0, 11025.0, 8000.0, 5512.0):
        n = int(fs); y = port.resample(tone(1000.0, n, fs), fs); m = np.arange(len(y))
        assert len(y) == int(np.floor(n * OUT / fs + 1e-9))
        assert np.abs(y[150:-150] - 0.5 * np.sin(2 * np.pi * 1000.0 * m[150:-150] / OUT)).max() < 2e-3, fs


def test_host_rules(lb):
    d = lb.Detective()
    assert d.recording_rate == 44100.0
    assert d.resampled_length(441000) == 55120
    assert d.set_recording_rate(48000.0) == 0 and d.recording_rate == 48000.0
    assert d.resampled_length(48000) == 5512
    assert d.set_recording_rate(0.0) == 1                                          # kLBAudioDetectiveArgumentInvalid
    assert d.set_recording_rate(4000.0) == 0 and d.resampled_length(4000) == 0     # below the processing rate: nothing to produce


@pytest.mark.gpu
@pytest.mark.parametrize("fs", [44100.0, 48000.0, 22050.0, 16000.0, 8000.0, 5512.0])
def test_kernel_equals_oracle_bit_for_bit(lb, port, fs):
    rng = np.random.default_rng(int(fs))
    n = int(3.3 * fs) + 17
    x = (0.4 * np.sin(2 * np.pi * 700.0 * np.arange(n) / fs) + 0.2 * rng.standard_normal(n)).astype(np.float32)
    d = lb.Detective(); d.set_recording_rate(fs)
    got = d.resample(x); want = port.resample(x, fs)
    assert got.shape == want.shape
    assert np.array_equal(got, want), (fs, np.abs(got - want).max())


@pytest.mark.gpu
def test_kernel_44k_tile_edges_and_long_clip(lb, port):
    """The 44.1 kHz kernel pairs outputs in tiles of 512: lengths around the tile and pair boundaries (odd tails, a single output),
    and a clip long enough for every irregular step (three instead of two stage-1 samples, once per 5,512 outputs) and every coarse-phase
    step (once per 86 outputs) to fall on each side of a pair."""
    d = lb.Detective()
    rng = np.random.default_rng(44)
    for n_out in (1, 2, 3, 511, 512, 513, 1023, 1024, 1025, 1537):
        n = int(np.ceil(n_out * FS / OUT)) + 1
        x = (0.3 * rng.standard_normal(n)).astype(np.float32)
        got = d.resample(x); want = port.resample(x, FS)
        assert len(want) in (n_out, n_out + 1) and np.array_equal(got, want), n_out
    x = (0.3 * rng.standard_normal(11 * 44100 + 5)).astype(np.float32)
    got = d.resample(x); want = port.resample(x, FS)
    assert len(want) > 60000 and np.array_equal(got, want)


@pytest.mark.gpu
def test_process_recorded_pcm_equals_process_of_resampled(lb, port):
    """One call from 44.1 kHz PCM = resample + LBAudioDetectiveProcessPCM, without the host round trip."""
    hi = port.synth_clip(21, 8 * 44100, 44100.0)
    d = lb.Detective()
    fp = d.process_recorded_pcm(hi)
    low = port.resample(hi, FS)
    want = d.process_pcm(low)
    assert fp.count == want.count == 5 and fp.equal(want)
    cfg = __import__("oracle.oracle", fromlist=["Cfg"]).Cfg.default()
    assert np.array_equal(fp.booleans(), port.process(cfg, low)) or (fp.booleans() != port.process(cfg, low)).mean() <= 1e-3


@pytest.mark.gpu
def test_process_recorded_batch_device(lb, port):
    import torch
    clips = np.stack([port.synth_clip(30 + c, 5 * 44100, 44100.0) for c in range(3)])
    d = lb.Detective()
    x = torch.from_numpy(clips).cuda(); n_out = d.resampled_length(clips.shape[1]); c = d.subfingerprints_for_length(n_out)
    w = torch.zeros((3, c, 8), dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        d.process_recorded_batch_device(x.data_ptr(), 3, clips.shape[1], clips.shape[1], w.data_ptr(), s.cuda_stream)
    s.synchronize()
    for i in range(3):
        assert np.array_equal(w[i].cpu().numpy().view(np.uint32), d.process_recorded_pcm(clips[i]).packed())
