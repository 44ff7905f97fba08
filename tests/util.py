"""Shared helpers of the parity tests: seeded signals and error metrics."""
import numpy as np


def rel_rms(a, b):
    """||a-b||_2 / ||b||_2  (RMS-normalised relative error; b is the oracle)."""
    a = np.asarray(a, np.float64).reshape(-1) if not np.iscomplexobj(a) else np.asarray(a, np.complex128).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1) if not np.iscomplexobj(b) else np.asarray(b, np.complex128).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def max_rel(a, b):
    """max |a-b| / max |b|"""
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))


def noise_iq(n, seed, amp=1.0):
    """i.i.d. uniform[-amp, amp) re/im -- the SpeedTester distribution (core/src/dsp/bench/speed_tester.h:37-41)."""
    rng = np.random.default_rng(seed)
    x = np.empty(2 * n, np.float32)
    x[:] = rng.uniform(-amp, amp, 2 * n)
    return x.view(np.complex64)


def fm_carrier(n, fs, offset, dev=75000.0, tones=((1000.0, 0.5), (5000.0, 0.3)), amp=0.05, start=0):
    """FM carrier at `offset` Hz: instantaneous frequency offset + dev * sum(a_i sin(2 pi f_i t))."""
    t = (np.arange(n, dtype=np.float64) + start) / fs
    phase = 2 * np.pi * offset * t
    for f, a in tones:
        phase += -(dev * a / f) * np.cos(2 * np.pi * f * t)
    return (amp * np.exp(1j * phase)).astype(np.complex64)


def am_carrier(n, fs, offset, tone=1000.0, depth=0.5, amp=0.05):
    t = np.arange(n, dtype=np.float64) / fs
    env = 1.0 + depth * np.sin(2 * np.pi * tone * t)
    return (amp * env * np.exp(2j * np.pi * offset * t)).astype(np.complex64)


def ssb_tone(n, fs, offset, tone=1000.0, amp=0.05):
    t = np.arange(n, dtype=np.float64) / fs
    return (amp * np.exp(2j * np.pi * (offset + tone) * t)).astype(np.complex64)


def test_signal(n, fs, carriers, seed=0x5D12, noise_amp=0.01):
    """noise + FM carriers at the given (offset, ...) list."""
    x = noise_iq(n, seed, noise_amp).copy()
    for c in carriers:
        x += fm_carrier(n, fs, c)
    return x


def to_i16(x):
    """complex64 in [-1,1) -> interleaved int16 IQ (file_source format)."""
    v = np.clip(np.round(x.view(np.float32) * 32768.0), -32768, 32767).astype(np.int16)
    return v
