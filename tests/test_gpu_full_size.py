"""GPU parity at the sizes and plans the headline numbers are quoted on (VERDICT r01 "weak" 1-3):

  * the bench configuration itself: BASELINE config 2, ONE 2^24-sample chunk, 8 WFM VFOs, default kernels -- every VFO and
    every FFT line against the oracle (fed in the reference's own <= 1e6-sample chunks; FM audio does not depend on the
    chunking, which test_chunking_invariance pins separately)
  * the 1.024 GS/s decimation plans of configs 4 / 5 (ratio 4096: first stage D = 64, T = 400; ratio 8192: D = 128, T = 726)
  * 16 WFM VFOs on one stream (config 5: more than one 16-VFO... exactly one full stage-1 / tail batch plus pairing)
  * RxVFO::setBandwidth through the fused front end with the radio AF chain behind the demodulator

Every comparison is CUDA-vs-oracle at the north-star tolerance 1e-5 (RMS-normalised), through the C ABI.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from util import rel_rms, noise_iq, fm_carrier

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def sb():
    import sdrplusplus_b200 as m
    from sdrplusplus_b200 import lib
    L = lib.load()
    assert L.b200_device_count() > 0
    assert L.b200_init(0) == 0
    return m


def _fm_into(x, fs, offset, amp=0.05, dev=75000.0, block=1 << 21):
    """adds an FM carrier in place, block by block (keeps the float64 temporaries small at 2^24 samples)."""
    for s in range(0, x.size, block):
        n = min(block, x.size - s)
        x[s:s + n] += fm_carrier(n, fs, offset, dev=dev, amp=amp, start=s)


def _oracle_vfo(oracle, xf, fs, cfg, chunk=1000000):
    """RxVFO -> WFM / NFM of one VFO, fed in reference-sized chunks (stream.h:9 caps a chunk at 1e6 samples)."""
    from sdrplusplus_b200 import lib as L
    v = oracle.rxvfo(fs, cfg.out_samplerate, cfg.bandwidth, cfg.offset)
    if cfg.demod == L.DEMOD_WFM:
        d = oracle.wfm(cfg.deviation, cfg.out_samplerate, False, cfg.low_pass)
    elif cfg.demod == L.DEMOD_NFM:
        d = oracle.nfm(cfg.out_samplerate, cfg.bandwidth, cfg.low_pass)
    else:
        d = None
    outs = []
    for i in range(0, xf.size // 2, chunk):
        y = v.process(xf[2 * i: 2 * (i + chunk)])
        outs.append(d.process(y) if d is not None else y)
    return np.concatenate(outs).reshape(-1, 2)


def _oracle_all(oracle, x, fs, cfgs, exact_phase=False):
    """exact_phase: the oracle's translator keeps an fp64 angle (same fp32 phaseDelta) instead of the reference's fp32
    recurrence `phase *= phaseDelta` -- BASELINE.md section 3 / SURVEY.md section 7: the recurrence's rounding walk is the
    REFERENCE's own numerical floor, it grows with the square root of the decimation and is reported, not gated."""
    xf = x.view(np.float32)
    oracle.set_rotator_mode(1 if exact_phase else 0)
    try:
        with ThreadPoolExecutor(max_workers=min(16, len(cfgs))) as ex:   # ctypes releases the GIL: one core per VFO
            return list(ex.map(lambda c: _oracle_vfo(oracle, xf, fs, c), cfgs))
    finally:
        oracle.set_rotator_mode(0)


def _gate_with_floor(outs, ids, ref, ref_exact, skip):
    """per VFO: (gpu vs faithful oracle, gpu vs exact-phase oracle, faithful vs exact-phase = the reference's floor)"""
    rows = []
    for vid, ya, yx in zip(ids, ref, ref_exact):
        assert outs[vid].shape == ya.shape == yx.shape
        rows.append((rel_rms(outs[vid][skip:], ya[skip:]), rel_rms(outs[vid][skip:], yx[skip:]), rel_rms(ya[skip:], yx[skip:])))
    return rows


def test_bench_configuration_one_16mi_chunk_vs_oracle(sb, oracle, report):
    """The configuration bench.py times, at the length BASELINE.md section 3 allows for parity (2^24 samples)."""
    fs, n = 100e6, 1 << 24
    offs = [5e6, -5e6, 15e6, -15e6, 25e6, -25e6, 35e6, -35e6]
    x = noise_iq(n, 0x5D12, 0.01).copy()
    for o in offs:
        _fm_into(x, fs, o)
    fe = sb.FrontEnd(fs, n)
    fe.set_fft(1 << 20, 20.0, 2)
    cfgs = [sb.VfoConfig.wfm(o) for o in offs]
    ids = [fe.add_vfo(c) for c in cfgs]
    tma0 = fe.stat("s1_tma_launches")
    outs, lines = fe.process(x)
    assert fe.stat("s1_tma_launches") == tma0 + 1            # the default stage-1 kernel of the bench really ran
    rows = _gate_with_floor(outs, ids, _oracle_all(oracle, x, fs, cfgs), _oracle_all(oracle, x, fs, cfgs, True), 1500)
    errs = [r[0] for r in rows]                  # config 2 is gated against the FAITHFUL oracle (fp32 phase recurrence)
    # spectrum branch: frames start every 5e6 samples; 4 complete inside the chunk
    skip, nz = oracle.fft_params(fs, 1 << 20, 20.0)
    la = np.array([oracle.fft_frame(1 << 20, nz, 2, x[f:f + nz]) for f in range(0, n - nz + 1, nz + skip)])
    assert lines.shape == la.shape == (4, 1 << 20)
    p, pr = 10.0 ** (lines.astype(np.float64) / 10), 10.0 ** (la.astype(np.float64) / 10)
    e_fft = float(np.max(np.abs(p - pr)) / np.max(pr))
    report["bench_config_2p24_one_chunk"] = {"wfm_audio_rel_rms": errs, "fft_power_rel_max": e_fft,
                                             "gpu_vs_exact_phase_oracle": [r[1] for r in rows],
                                             "reference_floor_faithful_vs_exact_phase": [r[2] for r in rows]}
    assert np.array_equal(np.argmax(lines, axis=1), np.argmax(la, axis=1))
    assert e_fft < TOL, e_fft
    assert max(errs) < TOL, errs
    fe.close()


@pytest.mark.parametrize("s1", [8, 7, 6])
def test_config5_16_wfm_vfos_1024msps_plan(sb, oracle, report, s1):
    """BASELINE config 5 per-GPU graph: 1.024 GS/s stream, 16 WFM VFOs (ratio-4096 plan (64,400)(8,36)(4,27)(2,69),
    plans.h:124-139), 1M-pt FFT; offsets on the 32 MHz grid so the filter-bank stage 1 applies."""
    fs, chunk, nch = 1.024e9, 1 << 22, 3
    n = chunk * nch
    offs = [(2 * k + 1) * 32e6 * sgn for k in range(8) for sgn in (1, -1)]
    x = noise_iq(n, 77, 0.01).copy()
    for o in offs:
        _fm_into(x, fs, o)
    fe = sb.FrontEnd(fs, chunk)
    fe.set_option("s1", s1)
    fe.set_fft(1 << 20, 20.0, 2)
    cfgs = [sb.VfoConfig.wfm(o) for o in offs]
    ids = [fe.add_vfo(c) for c in cfgs]
    outs, lines = fe.process_chunks(x, chunk)
    ref, ref_exact = _oracle_all(oracle, x, fs, cfgs), _oracle_all(oracle, x, fs, cfgs, True)
    assert ref[0].shape[0] == n // 4096
    rows = _gate_with_floor(outs, ids, ref, ref_exact, 600)
    report["c5_16vfo_1024msps_s1v%d" % s1] = {"gpu_vs_exact_phase_oracle": [r[1] for r in rows], "gpu_vs_faithful_oracle": [r[0] for r in rows],
                                              "reference_floor_faithful_vs_exact_phase": [r[2] for r in rows]}
    # 4096x decimation: the reference's fp32 phase recurrence walks 4e-5 ... 1e-4 of the audio away from its own exact-phase
    # form (measured, per offset).  Gate: the CUDA path is within 1e-5 of the exact-phase oracle, and no further from the
    # faithful oracle than that floor.
    for e_f, e_x, fl in rows:
        assert e_x < TOL, rows
        assert e_f < fl + TOL, rows
    fe.close()


def test_config4_plans_ratio_8192(sb, oracle, report):
    """1.024 GS/s -> NFM (50 kS/s): ratio-8192 plan (128,726)(8,36)(4,27)(2,69) + 2/5 polyphase resampler;
    offsets on config 4's 12.5 MHz grid (off every filter-bank grid: the per-VFO complex-tap stage 1 runs)."""
    fs, chunk, nch = 1.024e9, 1 << 22, 6
    n = chunk * nch
    offs = [12.5e6, -37.5e6, 112.5e6]
    x = noise_iq(n, 78, 0.01).copy()
    for o in offs:
        _fm_into(x, fs, o, dev=5000.0)
    fe = sb.FrontEnd(fs, chunk)
    cfgs = [sb.VfoConfig.nfm(o) for o in offs]
    ids = [fe.add_vfo(c) for c in cfgs]
    outs, _ = fe.process_chunks(x, chunk)
    ref, ref_exact = _oracle_all(oracle, x, fs, cfgs), _oracle_all(oracle, x, fs, cfgs, True)
    rows = _gate_with_floor(outs, ids, ref, ref_exact, 500)
    report["c4_nfm_1024msps_ratio8192"] = {"gpu_vs_exact_phase_oracle": [r[1] for r in rows], "gpu_vs_faithful_oracle": [r[0] for r in rows],
                                           "reference_floor_faithful_vs_exact_phase": [r[2] for r in rows], "outputs": int(ref[0].shape[0])}
    assert ref[0].shape[0] > 900
    for e_f, e_x, fl in rows:
        assert e_x < TOL, rows
        assert e_f < fl + TOL, rows
    fe.close()


def test_set_bandwidth_with_af_chain_and_bypass(sb, oracle, report):
    """RxVFO::setBandwidth (rx_vfo.h:60-70) must retune the CHANNEL filter, not a stage of the AF chain behind the
    demodulator; bandwidth == outSamplerate bypasses the filter and a later change brings it back."""
    FS = 2.4e6
    n, chunk = 600000, 12000
    x = noise_iq(n, 31, 0.01).copy() + fm_carrier(n, FS, 300e3)
    fe = sb.FrontEnd(FS, chunk)
    cfg = sb.VfoConfig.wfm(300e3).with_af(48000.0, high_pass=True, deemph_tau=50e-6)
    vid = fe.add_vfo(cfg)
    cfg2 = sb.VfoConfig.raw(300e3, 250e3, 250e3)                      # created bypassed (bandwidth == outSR)
    vid2 = fe.add_vfo(cfg2)
    v, d = oracle.rxvfo(FS, 250e3, 150e3, 300e3), oracle.wfm(75e3, 250e3)
    af = oracle.resamp_stereo(250e3, 48000.0)
    hp = oracle.fir_cr(oracle.highpass(300.0, 100.0, 48000.0))
    de = oracle.deemph(50e-6, 48000.0)
    v2 = oracle.rxvfo(FS, 250e3, 250e3, 300e3)
    xf = x.view(np.float32)
    yg, ya, yg2, ya2 = [], [], [], []
    for k, i in enumerate(range(0, n, chunk)):
        if k == 12:
            fe.set_vfo_bandwidth(vid, 100e3); v.set_bandwidth(100e3)
            fe.set_vfo_bandwidth(vid2, 120e3); v2.set_bandwidth(120e3)          # bypass -> low-pass
        if k == 30:
            fe.set_vfo_bandwidth(vid, 180e3); v.set_bandwidth(180e3)
            fe.set_vfo_bandwidth(vid2, 250e3); v2.set_bandwidth(250e3)          # low-pass -> bypass
        outs, _ = fe.process(x[i:i + chunk])
        yg.append(outs[vid]); yg2.append(outs[vid2])
        seg = xf[2 * i: 2 * (i + chunk)]
        ya.append(de.process(hp.process(af.process(d.process(v.process(seg))))).reshape(-1, 2))
        ya2.append(v2.process(seg).view(np.complex64))
    yg, ya, yg2, ya2 = np.concatenate(yg), np.concatenate(ya), np.concatenate(yg2), np.concatenate(ya2)
    assert yg.shape == ya.shape and yg2.shape == ya2.shape
    e1, e2 = rel_rms(yg[2500:], ya[2500:]), rel_rms(yg2[2000:], ya2[2000:])
    report["set_bandwidth_af_chain"] = {"wfm_af_rel_rms": e1, "raw_bypass_rel_rms": e2}
    assert e1 < TOL, e1
    assert e2 < TOL, e2
    fe.close()
