"""GPU parity of the optional branches of the demodulators: the stereo branch of
demod::BroadcastFM (pilot band-pass -> loop::PLL -> L-R recovery -> two audio low-passes, broadcast_fm.h:147-190) and
noise_reduction::PowerSquelch (power_squelch.h:33-50), as stand-alone blocks and inside the fused front end."""
import numpy as np
import pytest

from golden_cases import stereo_mpx_iq
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


@pytest.mark.parametrize("low_pass,chunk", [(True, 1250), (False, 3001)])
def test_wfm_stereo_block(sb, oracle, report, low_pass, chunk):
    fs, n = 250e3, 150000
    x = stereo_mpx_iq(n, fs).view(np.float32)
    y = sb.Block.wfm(75e3, fs, stereo=True, lowpass=low_pass).process_chunks(x, chunk).reshape(-1, 2)
    ya = oracle.wfm(75e3, fs, True, low_pass).process_chunks(x, chunk).reshape(-1, 2)
    assert y.shape == ya.shape
    # the PLL pulls in over the first few thousand samples; once locked the loop is stable and both follow the same pilot
    e = rel_rms(y[20000:], ya[20000:])
    e_lock = rel_rms(y[:20000], ya[:20000])
    sep = float(np.std(y[20000:, 0] - y[20000:, 1]))                # the L-R programme really is there (0.3 * 2 / sqrt 2)
    report["wfm_stereo_block_lp%d" % int(low_pass)] = {"locked_rel_rms": e, "pull_in_rel_rms": e_lock, "l_minus_r_std": sep}
    assert e < TOL, e
    assert e_lock < 1e-4, e_lock
    if low_pass:
        assert 0.3 < sep < 0.6, sep


def test_wfm_stereo_in_front_end(sb, oracle, report):
    """2.4 MS/s stream, one stereo WFM VFO next to a mono one: RxVFO -> BroadcastFM(stereo) per chunk."""
    FS, n, chunk = 2.4e6, 960000, 24000
    base = stereo_mpx_iq(n // 8 + 16, 300e3)                       # build the carrier at 300 kS/s, then place it at +300 kHz
    up = np.repeat(base, 8)[:n]                                    # zero-order hold is enough for a parity signal
    t = np.arange(n) / FS
    x = (0.3 * up * np.exp(2j * np.pi * 300e3 * t)).astype(np.complex64) + noise_iq(n, 9, 0.005) + fm_carrier(n, FS, -650e3)
    fe = sb.FrontEnd(FS, chunk)
    cfg_s, cfg_m = sb.VfoConfig.wfm_stereo(300e3), sb.VfoConfig.wfm(-650e3)
    vs, vm = fe.add_vfo(cfg_s), fe.add_vfo(cfg_m)
    outs, _ = fe.process_chunks(x, chunk)
    v, d = oracle.rxvfo(FS, 250e3, 150e3, 300e3), oracle.wfm(75e3, 250e3, True, True)
    v2, d2 = oracle.rxvfo(FS, 250e3, 150e3, -650e3), oracle.wfm(75e3, 250e3)
    ya, yb = [], []
    xf = x.view(np.float32)
    for i in range(0, n, chunk):
        seg = xf[2 * i: 2 * (i + chunk)]
        ya.append(d.process(v.process(seg)).reshape(-1, 2))
        yb.append(d2.process(v2.process(seg)).reshape(-1, 2))
    ya, yb = np.concatenate(ya), np.concatenate(yb)
    assert outs[vs].shape == ya.shape and outs[vm].shape == yb.shape
    e_s, e_m = rel_rms(outs[vs][30000:], ya[30000:]), rel_rms(outs[vm][4000:], yb[4000:])
    report["frontend_wfm_stereo"] = {"stereo_rel_rms": e_s, "mono_neighbour_rel_rms": e_m}
    assert e_m < TOL, e_m
    assert e_s < TOL, e_s
    fe.close()


def test_power_squelch_block_and_front_end(sb, oracle, report):
    FS, n, chunk = 2.4e6, 480000, 12000
    env = np.concatenate([np.full(n // 4, 1.0), np.full(n // 4, 0.02), np.full(n // 4, 0.5), np.full(n - 3 * (n // 4), 0.01)]).astype(np.float32)
    x = ((fm_carrier(n, FS, 300e3) * env) + noise_iq(n, 4, 0.0005)).astype(np.complex64)
    # stand-alone block at an IF rate: levels well inside and outside the threshold
    xi = (noise_iq(60000, 6, 1.0) * np.repeat(np.array([0.5, 0.001, 0.2, 0.0005, 0.05, 0.3], np.float32), 10000)).astype(np.complex64)
    yb = sb.Block.squelch(-30.0).process_chunks(xi.view(np.float32), 2500)
    ya = oracle.squelch(-30.0).process_chunks(xi.view(np.float32), 2500)
    assert np.array_equal(yb.view(np.uint32), ya.view(np.uint32))                 # copy or zero: bit for bit
    assert 0 < np.count_nonzero(ya) < ya.size
    # in the front end: RxVFO -> PowerSquelch -> WFM (radio IF chain, radio_module.h:88-96)
    fe = sb.FrontEnd(FS, chunk)
    vid = fe.add_vfo(sb.VfoConfig.wfm(300e3).with_squelch(-40.0))
    outs, _ = fe.process_chunks(x, chunk)
    v, q, d = oracle.rxvfo(FS, 250e3, 150e3, 300e3), oracle.squelch(-40.0), oracle.wfm(75e3, 250e3)
    xf = x.view(np.float32)
    ref = np.concatenate([d.process(q.process(v.process(xf[2 * i: 2 * (i + chunk)]))).reshape(-1, 2) for i in range(0, n, chunk)])
    assert outs[vid].shape == ref.shape
    e = rel_rms(outs[vid][4000:], ref[4000:])
    muted = float(np.mean(np.abs(ref[n // 4 // 10 + 2000: n // 2 // 10 - 2000])))      # the 0.02-amplitude quarter is closed
    report["frontend_power_squelch"] = {"wfm_audio_rel_rms": e, "muted_mean_abs": muted}
    assert e < TOL, e
    fe.close()
