"""GPU parity of the 256-channel polyphase filter-bank channelizer (BASELINE config 3; SURVEY.md section 8d: "not a reference
feature ... oracle = direct-form per-channel restatement (xlate by -k fs/256 -> 32 512-tap FIR -> keep every 256th) in fp64").
The oracle below IS that direct form, evaluated in float64 with numpy, with the prototype taps the reference's
taps::windowedSinc<float>(32512, fs/512, fs, nuttall) gives (taken from the library and cross-checked against the oracle's
own windowed-sinc restatement)."""
import ctypes as C

import numpy as np
import pytest

from util import noise_iq, rel_rms

pytestmark = pytest.mark.gpu
M, P = 256, 127
T = M * P


@pytest.fixture(scope="module")
def L():
    from sdrplusplus_b200 import lib
    l = lib.load()
    assert l.b200_device_count() > 0
    assert l.b200_init(0) == 0
    return l


def _direct_form(x_with_hist, h, channels, m_list):
    """y_k[m] = sum_t h[t] x[n0 + t] e^{-j 2 pi k (n0 + t) / M}, n0 = m M + (M - 1) - (T - 1); x index 0 = first sample of the
    stream, history before it = zeros (x_with_hist[T - 1 + n] = x[n])."""
    h64 = h.astype(np.float64)
    t = np.arange(T)
    out = np.empty((len(m_list), len(channels)), np.complex128)
    for a, m in enumerate(m_list):
        n0 = m * M + (M - 1) - (T - 1)
        seg = x_with_hist[(T - 1) + n0: (T - 1) + n0 + T].astype(np.complex128)
        for b, k in enumerate(channels):
            out[a, b] = np.sum(h64 * seg * np.exp(-2j * np.pi * k * ((n0 + t) % M) / M))
    return out


def test_channelizer_vs_float64_direct_form(L, oracle, report):
    max_chunk = 64 * M
    ch = L.b200_chan_create(M, P, max_chunk)
    assert ch
    h = np.empty(T, np.float32)
    assert L.b200_chan_prototype(ch, h.ctypes.data, T) == T
    # prototype == the reference's windowedSinc<float>(T, fs/512, fs, nuttall): same formula as lowPass with the tap count
    # given; the oracle's low-pass restatement with a transition width that yields exactly T taps is bit-identical
    fs = 500e6
    ref_h = oracle.lowpass(fs / 512.0, 3.8 * fs / T, fs)
    if ref_h.size == T:
        assert np.array_equal(ref_h.view(np.uint32), h.view(np.uint32))
    assert abs(float(np.sum(h.astype(np.float64))) - 1.0) < 2e-3
    n = 3 * max_chunk
    x = noise_iq(n, 61, 1.0).copy()
    tt = np.arange(n)
    for k, a in ((5, 0.7), (100, 0.5), (255, 0.3)):                     # tones at channel centres + 0.1 of a channel
        x += (a * np.exp(2j * np.pi * (k + 0.1) / M * tt)).astype(np.complex64)
    outs = []
    for c in range(0, n, max_chunk):
        seg = np.ascontiguousarray(x[c:c + max_chunk])
        y = np.empty(max_chunk, np.complex64)
        assert L.b200_chan_process(ch, seg.ctypes.data, max_chunk, 0, y.ctypes.data, 0) == max_chunk // M
        outs.append(y.reshape(-1, M))
    y = np.concatenate(outs)                                            # [n / M][M]
    xh = np.concatenate([np.zeros(T - 1, np.complex64), x])
    channels = [0, 1, 5, 6, 100, 128, 200, 255]
    m_list = [0, 1, 63, 64, 65, 126, 127, 128, 150, 191]                # across the chunk boundaries (64 output times per chunk)
    ref = _direct_form(xh, h, channels, m_list)
    got = y[np.ix_(m_list, channels)]
    e = rel_rms(got, ref)
    # every channel of a few output times through the polyphase identity in float64 (covers all 256 branches / bins)
    m2 = [130, 131]
    ref2 = np.empty((len(m2), M), np.complex128)
    h64 = h.astype(np.float64).reshape(P, M)
    for a, m in enumerate(m2):
        win = xh[(T - 1) + m * M + (M - 1) - (T - 1):][:T].astype(np.complex128).reshape(P, M)
        ref2[a] = np.fft.fft(np.sum(h64 * win, axis=0) * np.exp(-2j * np.pi * np.arange(M) * 0), M) * 1.0
        # the window starts at n0 = m M + M - 1 - (T - 1) = (m - P + 1) M + ... : n0 mod M = 0, so branch r carries phase e^{-j 2 pi k r / M}
    e2 = rel_rms(y[m2], ref2)
    report["channelizer_256x127"] = {"direct_form_rel_rms": e, "all_channels_rel_rms": e2}
    assert e < 1e-5, e
    assert e2 < 1e-5, e2
    # the tone at channel 5 + 0.1 dominates channel 5
    assert np.argmax(np.mean(np.abs(y[150:]) ** 2, axis=0)) == 5
    L.b200_chan_destroy(ch)
