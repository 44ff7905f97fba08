"""smoke(): one small invocation of the hot path on cuda:0 through the C ABI, checked against the CPU oracle."""
import os
import sys

import numpy as np


def smoke():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import sdrplusplus_b200 as sb
    from sdrplusplus_b200 import lib
    from oracle.oracle import Oracle          # the checker, never the thing measured

    L = lib.load()
    if L.b200_device_count() <= 0:
        raise RuntimeError("smoke(): no CUDA device -- libb200dsp has no CPU fallback")
    lib.check(L.b200_init(0))
    fs, n, chunk = 2.4e6, 240000, 12000
    rng = np.random.default_rng(0x5D12)
    x = (rng.uniform(-0.02, 0.02, n) + 1j * rng.uniform(-0.02, 0.02, n)).astype(np.complex64)
    t = np.arange(n) / fs
    offs = (300e3, -650e3)
    for o in offs:
        x += (0.05 * np.exp(1j * (2 * np.pi * o * t - (75e3 * 0.5 / 1e3) * np.cos(2 * np.pi * 1e3 * t)))).astype(np.complex64)

    fe = sb.FrontEnd(fs, chunk)
    fe.set_fft(65536, 20.0, lib.WIN_NUTTALL)
    cfgs = [sb.VfoConfig.wfm(o) for o in offs]
    ids = [fe.add_vfo(c) for c in cfgs]
    outs, lines = fe.process_chunks(x, chunk)
    launches = fe.launch_count()
    fe.close()

    orc = Oracle("restatement")
    worst = 0.0
    for vid, c in zip(ids, cfgs):
        v = orc.rxvfo(fs, c.out_samplerate, c.bandwidth, c.offset)
        d = orc.wfm(c.deviation, c.out_samplerate, False, True)
        ref = np.concatenate([d.process(v.process(x.view(np.float32)[2 * i: 2 * (i + chunk)])) for i in range(0, n, chunk)]).reshape(-1, 2)
        got = outs[vid]
        assert got.shape == ref.shape, (got.shape, ref.shape)
        e = float(np.linalg.norm(got[4000:] - ref[4000:]) / np.linalg.norm(ref[4000:]))
        worst = max(worst, e)
    skip, nz = orc.fft_params(fs, 65536, 20.0)
    assert lines.shape == (2, 65536), lines.shape
    ref_line = orc.fft_frame(65536, nz, 2, x[:nz])
    p, pr = 10.0 ** (lines[0].astype(np.float64) / 10), 10.0 ** (ref_line.astype(np.float64) / 10)
    e_fft = float(np.max(np.abs(p - pr)) / np.max(pr))
    assert int(np.argmax(lines[0])) == int(np.argmax(ref_line))
    assert worst < 1e-5, "WFM audio differs from the oracle: %g" % worst
    assert e_fft < 1e-5, "FFT line differs from the oracle: %g" % e_fft
    print("smoke ok: 2 VFO WFM rel-rms %.2e, FFT power rel-max %.2e, %d kernel launches" % (worst, e_fft, launches))
