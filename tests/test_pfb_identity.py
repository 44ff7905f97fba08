"""The stage-1 filter-bank identity (csrc/xd_pfb.cuh) checked in float64 on the CPU:

    y_v[m] = e^{j phi_v(i_m)} sum_k h[k] e^{j w_v k} x[i_m + k]
           = e^{j phi_v(i_m)} sum_{a < PS} e^{j w_v a} S_a[m],   S_a[m] = sum_{k = a (PS)} sigma^floor(k/PS) h[k] x[i_m + k]

with w_v the angle of the fp32-rounded phasor the reference rotates by (frequency_xlator.h:17), which sits a few
1e-8 rad off the frequency grid: the test measures what that costs and holds it to the bound the scheduler accepts."""
import numpy as np
import pytest

FS = 100e6
OFFSETS = [5e6, -5e6, 15e6, -15e6, 25e6, -25e6, 35e6, -35e6]          # BASELINE config 2


def _w_eff(offset_hz, fs):
    """angle per sample the reference applies: angle of ((float)cos w, (float)sin w), w = 2 pi (-offset) / fs"""
    w = 2.0 * np.pi * (-offset_hz) / fs
    return float(np.arctan2(np.float64(np.float32(np.sin(w))), np.float64(np.float32(np.cos(w)))))


def _detect(ws, T, tol_rad=8e-6):
    """the scheduler's acceptance test (engine.cpp, Scheduler::run): smallest P <= 10 with e^{j w P} = +-1 for every VFO"""
    for P in range(1, 11):
        tol = tol_rad * P / T
        sg = 0
        ok = True
        for w in ws:
            r = (w * P) % (2 * np.pi)
            d0 = min(r, 2 * np.pi - r)
            d1 = abs(r - np.pi)
            sv = 1 if d0 <= tol else (-1 if d1 <= tol else 0)
            if sv == 0 or (sg and sv != sg):
                ok = False
                break
            sg = sv
        if ok:
            for PS in (8, 10):
                if PS % P == 0:
                    return PS, (-1 if (sg < 0 and ((PS // P) & 1)) else 1)
            return 0, 1
    return 0, 1


def test_config2_plan_is_a_filter_bank_and_off_grid_plans_are_not():
    ws = [_w_eff(o, FS) for o in OFFSETS]
    assert _detect(ws, 143) == (10, -1)
    assert _detect([_w_eff(300e3, 2.4e6)], 27) == (8, 1)                      # config 1: 1/8 of the sample rate
    assert _detect([_w_eff(o, FS) for o in (5e6, -7e6, 15e6, -17e6)], 143)[0] == 0
    assert _detect([_w_eff(5e6 + 40.0, FS)], 143)[0] == 0                     # 40 Hz off the grid: 3.6e-4 rad over the taps


def test_filter_bank_form_equals_the_per_vfo_form(oracle, report):
    h = np.asarray(oracle.decim_taps(256, 0), np.float64)                     # first stage of the 100 MS/s -> 390 kS/s plan
    T, D, PS, sigma = h.size, 32, 10, -1
    assert T == 143
    rng = np.random.default_rng(11)
    n_out = 400
    x = (rng.uniform(-1, 1, n_out * D + T) + 1j * rng.uniform(-1, 1, n_out * D + T)).astype(np.complex64).astype(np.complex128)
    k = np.arange(T)
    a = k % PS
    sgn = np.where((k // PS) % 2, float(sigma), 1.0)
    worst = 0.0
    for off in OFFSETS:
        w = _w_eff(off, FS)
        i_m = np.arange(n_out) * D
        win = x[i_m[:, None] + k[None, :]]                                    # [n_out, T]
        y_ref = np.exp(1j * w * i_m) * (win * (h * np.exp(1j * w * k))[None, :]).sum(axis=1)
        S = np.zeros((n_out, PS), np.complex128)
        for aa in range(PS):
            sel = a == aa
            S[:, aa] = (win[:, sel] * (h[sel] * sgn[sel])[None, :]).sum(axis=1)
        eps = ((w * PS - np.pi + np.pi) % (2 * np.pi)) - np.pi                # signed distance of w*PS from pi (sigma = -1)
        corr = np.exp(1j * eps * (T // (2 * PS)))                             # drift centred on the middle of the window
        y_pfb = np.exp(1j * w * i_m) * corr * (S * np.exp(1j * w * np.arange(PS))[None, :]).sum(axis=1)
        err = float(np.max(np.abs(y_pfb - y_ref)) / np.max(np.abs(y_ref)))
        worst = max(worst, err)
    report["pfb_identity_float64_worst_rel"] = worst
    assert worst < 1e-6, worst
    # exactly on the grid the identity is exact to rounding
    w = 2.0 * np.pi * 3 / 20
    win = x[(np.arange(n_out) * D)[:, None] + k[None, :]]
    y_ref = (win * (h * np.exp(1j * w * k))[None, :]).sum(axis=1)
    S = np.stack([(win[:, a == aa] * (h[a == aa] * sgn[a == aa])[None, :]).sum(axis=1) for aa in range(PS)], axis=1)
    y_pfb = (S * np.exp(1j * w * np.arange(PS))[None, :]).sum(axis=1)
    assert np.max(np.abs(y_pfb - y_ref)) / np.max(np.abs(y_ref)) < 1e-13
