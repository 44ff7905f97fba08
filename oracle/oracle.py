"""ctypes binding of oracle/sdrpp_oracle.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two interchangeable implementations of the same C API:

* ``Oracle("restatement")`` -> oracle/liboracle.so  (plain-C restatement, travels with the repo)
* ``Oracle("reference")``   -> oracle/_ref/libsdrpp_ref.so (the reference's own dsp headers over the
  restated VOLK/FFT leaf layer; built only where /root/reference exists, shipped prebuilt to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_PATHS = {
    "restatement": os.path.join(_HERE, "liboracle.so"),
    "reference": os.path.join(_HERE, "_ref", "libsdrpp_ref.so"),
    "restatement_fast": os.path.join(_HERE, "liboracle_fast.so"),
    "reference_fast": os.path.join(_HERE, "_ref", "libsdrpp_ref_fast.so"),
}


class ResampPlan(C.Structure):
    _fields_ = [("mode", C.c_int), ("predec_ratio", C.c_int), ("interp", C.c_int), ("decim", C.c_int),
                ("ntaps", C.c_int), ("taps_per_phase", C.c_int)]


def available(kind):
    return os.path.exists(_PATHS[kind])


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Block:
    """One streaming reference block (state carried across process() calls)."""

    def __init__(self, lib, handle, in_floats, out_floats, max_ratio=1.0):
        self._lib, self._h = lib, handle
        self._in_f, self._out_f, self._ratio = in_floats, out_floats, max_ratio
        if not handle:
            raise RuntimeError("oracle block creation failed")

    def process(self, x):
        x = _f32(x).reshape(-1)
        count = x.size // self._in_f
        cap = int(count * self._ratio) + 64
        out = np.empty(cap * self._out_f, dtype=np.float32)
        n = self._lib.orc_process(self._h, count, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        if n < 0:
            raise RuntimeError("oracle process failed")
        return out[: n * self._out_f].copy()

    def process_chunks(self, x, chunk):
        """Feed x in chunks of `chunk` input samples (the reference caps a chunk at 1e6 samples)."""
        x = _f32(x).reshape(-1)
        step = chunk * self._in_f
        outs = [self.process(x[i:i + step]) for i in range(0, x.size, step)]
        return np.concatenate(outs) if outs else np.empty(0, np.float32)

    def reset(self):
        self._lib.orc_reset(self._h)

    def close(self):
        if self._h:
            self._lib.orc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RdsDemod:
    def __init__(self, lib):
        self._lib, self._h = lib, lib.orc_rdsdemod_create()
        if not self._h:
            raise RuntimeError("oracle block creation failed")

    def process(self, x):
        x = np.ascontiguousarray(x, np.complex64).reshape(-1)
        soft = np.empty(x.size + 16, np.float32)
        hard = np.empty(x.size + 16, np.uint8)
        n = self._lib.orc_rdsdemod_process(self._h, int(x.size), x.ctypes.data_as(C.c_void_p), soft.ctypes.data_as(C.c_void_p),
                                           hard.ctypes.data_as(C.c_void_p))
        if n < 0:
            raise RuntimeError("oracle process failed")
        return soft[:n].copy(), hard[:n].copy()

    def process_chunks(self, x, chunk):
        x = np.ascontiguousarray(x, np.complex64).reshape(-1)
        parts = [self.process(x[i:i + chunk]) for i in range(0, x.size, chunk)]
        if not parts:
            return np.empty(0, np.float32), np.empty(0, np.uint8)
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def reset(self):
        self._lib.orc_rdsdemod_reset(self._h)

    def close(self):
        if self._h:
            self._lib.orc_rdsdemod_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Oracle:
    def __init__(self, kind="restatement"):
        path = _PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        self.kind = kind
        L = self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        vp, d, i, fp = C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        sig = {
            "orc_impl": (C.c_char_p, []),
            "orc_set_rotator_mode": (None, [i]),
            "orc_estimate_tap_count": (i, [d, d]),
            "orc_lowpass": (i, [d, d, d, i, vp, i]),
            "orc_highpass": (i, [d, d, d, i, vp, i]),
            "orc_bandpass_c": (i, [d, d, d, d, i, vp, i]),
            "orc_window": (d, [i, d, d]),
            "orc_decim_plan": (i, [i, ip, ip, i]),
            "orc_decim_taps": (i, [i, i, vp, i]),
            "orc_resamp_plan_get": (i, [d, d, C.POINTER(ResampPlan)]),
            "orc_resamp_taps": (i, [d, d, vp, i]),
            "orc_xlator_create": (vp, [d, d]),
            "orc_xlator_set_offset": (None, [vp, d, d]),
            "orc_xlator_get_phase": (None, [vp, vp, vp]),
            "orc_decim_create": (vp, [i]),
            "orc_resamp_create": (vp, [d, d]),
            "orc_resamp_stereo_create": (vp, [d, d]),
            "orc_fir_cr_create": (vp, [vp, i]),
            "orc_fir_rr_create": (vp, [vp, i]),
            "orc_decfir_cr_create": (vp, [vp, i, i]),
            "orc_rxvfo_create": (vp, [d, d, d, d]),
            "orc_rxvfo_set_offset": (None, [vp, d]),
            "orc_rxvfo_set_bandwidth": (None, [vp, d]),
            "orc_quad_create": (vp, [d, d]),
            "orc_wfm_create": (vp, [d, d, i, i]),
            "orc_wfm_rds_create": (vp, [d, d]),
            "orc_nfm_create": (vp, [d, d, i]),
            "orc_am_create": (vp, [i, d, d, d, d, d]),
            "orc_ssb_create": (vp, [i, d, d, d, d]),
            "orc_dcblock_c_create": (vp, [d]),
            "orc_squelch_create": (vp, [d]),
            "orc_nb_create": (vp, [d, d]),
            "orc_fmif_create": (vp, [C.c_int]),
            "orc_deemph_create": (vp, [d, d]),
            "orc_rdsdemod_create": (vp, []),
            "orc_rdsdemod_process": (i, [vp, i, vp, vp, vp]),
            "orc_rdsdemod_reset": (None, [vp]),
            "orc_rdsdemod_free": (None, [vp]),
            "orc_rdsdemod_taps": (i, [vp, i, vp]),
            "orc_process": (i, [vp, i, vp, vp]),
            "orc_reset": (None, [vp]),
            "orc_free": (None, [vp]),
            "orc_fft_params": (None, [d, i, d, ip, ip]),
            "orc_window_buf": (None, [i, i, vp]),
            "orc_fft_create": (vp, [i, i, i]),
            "orc_fft_frame": (i, [vp, vp, vp]),
            "orc_fft_raw": (i, [vp, vp, vp]),
            "orc_fft_free": (None, [vp]),
            "orc_zoom": (None, [i, i, i, i, vp, vp]),
            "orc_hold": (None, [vp, vp, i, C.c_float]),
            "orc_i16_to_f32": (None, [vp, vp, i]),
            "orc_pcm_compress": (i, [i, i, vp, vp]),
            "orc_pcm_decompress": (i, [i, vp, vp]),
            "orc_export_convert": (None, [vp, i, i, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    def impl(self):
        return self.lib.orc_impl().decode()

    def set_rotator_mode(self, mode):
        """0 = faithful fp32 recurrence (declared oracle), 1 = exact-phase rotator (restatement only)."""
        self.lib.orc_set_rotator_mode(int(mode))

    # ---- design ----
    def estimate_tap_count(self, tw, sr):
        return self.lib.orc_estimate_tap_count(tw, sr)

    def lowpass(self, cutoff, tw, sr, odd=False):
        n = self.lib.orc_lowpass(cutoff, tw, sr, int(odd), None, 0)
        out = np.empty(n, np.float32)
        self.lib.orc_lowpass(cutoff, tw, sr, int(odd), out.ctypes.data_as(C.c_void_p), n)
        return out

    def highpass(self, cutoff, tw, sr, odd=False):
        n = self.lib.orc_highpass(cutoff, tw, sr, int(odd), None, 0)
        out = np.empty(n, np.float32)
        self.lib.orc_highpass(cutoff, tw, sr, int(odd), out.ctypes.data_as(C.c_void_p), n)
        return out

    def bandpass_c(self, b0, b1, tw, sr, odd=False):
        n = self.lib.orc_bandpass_c(b0, b1, tw, sr, int(odd), None, 0)
        out = np.empty(2 * n, np.float32)
        self.lib.orc_bandpass_c(b0, b1, tw, sr, int(odd), out.ctypes.data_as(C.c_void_p), n)
        return out.view(np.complex64)

    def window(self, typ, n, N):
        return self.lib.orc_window(typ, n, N)

    def decim_plan(self, ratio):
        D = (C.c_int * 8)()
        T = (C.c_int * 8)()
        ns = self.lib.orc_decim_plan(ratio, D, T, 8)
        return [(D[k], T[k]) for k in range(ns)]

    def decim_taps(self, ratio, stage):
        n = self.lib.orc_decim_taps(ratio, stage, None, 0)
        out = np.empty(n, np.float32)
        self.lib.orc_decim_taps(ratio, stage, out.ctypes.data_as(C.c_void_p), n)
        return out

    def resamp_plan(self, in_sr, out_sr):
        p = ResampPlan()
        self.lib.orc_resamp_plan_get(in_sr, out_sr, C.byref(p))
        return {k: getattr(p, k) for k, _ in ResampPlan._fields_}

    def resamp_taps(self, in_sr, out_sr):
        n = self.lib.orc_resamp_taps(in_sr, out_sr, None, 0)
        out = np.empty(n, np.float32)
        if n:
            self.lib.orc_resamp_taps(in_sr, out_sr, out.ctypes.data_as(C.c_void_p), n)
        return out

    # ---- blocks ----
    def xlator(self, offset_hz, sr):
        return Block(self.lib, self.lib.orc_xlator_create(offset_hz, sr), 2, 2)

    def xlator_phase(self, blk):
        ph = np.zeros(2, np.float32)
        dl = np.zeros(2, np.float32)
        self.lib.orc_xlator_get_phase(blk._h, ph.ctypes.data_as(C.c_void_p), dl.ctypes.data_as(C.c_void_p))
        return ph, dl

    def decim(self, ratio):
        return Block(self.lib, self.lib.orc_decim_create(ratio), 2, 2)

    def resamp(self, in_sr, out_sr):
        return Block(self.lib, self.lib.orc_resamp_create(in_sr, out_sr), 2, 2, max(1.0, out_sr / in_sr) * 1.01)

    def resamp_stereo(self, in_sr, out_sr):
        return Block(self.lib, self.lib.orc_resamp_stereo_create(in_sr, out_sr), 2, 2, max(1.0, out_sr / in_sr) * 1.01)

    def fir_cr(self, taps):
        t = _f32(taps)
        return Block(self.lib, self.lib.orc_fir_cr_create(t.ctypes.data_as(C.c_void_p), t.size), 2, 2)

    def fir_rr(self, taps):
        t = _f32(taps)
        return Block(self.lib, self.lib.orc_fir_rr_create(t.ctypes.data_as(C.c_void_p), t.size), 1, 1)

    def decfir_cr(self, taps, decim):
        t = _f32(taps)
        return Block(self.lib, self.lib.orc_decfir_cr_create(t.ctypes.data_as(C.c_void_p), t.size, decim), 2, 2)

    def rxvfo(self, in_sr, out_sr, bw, offset):
        b = Block(self.lib, self.lib.orc_rxvfo_create(in_sr, out_sr, bw, offset), 2, 2, max(1.0, out_sr / in_sr) * 1.01)
        b.set_offset = lambda off: self.lib.orc_rxvfo_set_offset(b._h, off)
        b.set_bandwidth = lambda bw_: self.lib.orc_rxvfo_set_bandwidth(b._h, bw_)
        return b

    def quad(self, dev, sr):
        return Block(self.lib, self.lib.orc_quad_create(dev, sr), 2, 1)

    def wfm(self, dev, sr, stereo=False, lowpass=True):
        return Block(self.lib, self.lib.orc_wfm_create(dev, sr, int(stereo), int(lowpass)), 2, 2)

    def nfm(self, sr, bw, lowpass=True):
        return Block(self.lib, self.lib.orc_nfm_create(sr, bw, int(lowpass)), 2, 2)

    def am(self, agc_mode, bw, attack, decay, dcrate, sr):
        return Block(self.lib, self.lib.orc_am_create(agc_mode, bw, attack, decay, dcrate, sr), 2, 2)

    def ssb(self, mode, bw, sr, attack, decay):
        return Block(self.lib, self.lib.orc_ssb_create(mode, bw, sr, attack, decay), 2, 2)

    def dcblock_c(self, rate):
        return Block(self.lib, self.lib.orc_dcblock_c_create(rate), 2, 2)

    def wfm_rds(self, dev, sr):
        """BroadcastFM's RDS side output: complex in at sr, complex out at 5 kS/s."""
        return Block(self.lib, self.lib.orc_wfm_rds_create(dev, sr), 2, 2)

    def squelch(self, level):
        return Block(self.lib, self.lib.orc_squelch_create(level), 2, 2)

    def noise_blanker(self, rate, level):
        return Block(self.lib, self.lib.orc_nb_create(rate, level), 2, 2)

    def fm_if(self, bins):
        return Block(self.lib, self.lib.orc_fmif_create(bins), 2, 2)

    def deemph(self, tau, sr):
        return Block(self.lib, self.lib.orc_deemph_create(tau, sr), 2, 2)

    def rds_demod(self):
        """RDSDemod (decoder_modules/radio/src/rds_demod.h): complex at 5 kS/s in -> (soft float32, hard uint8) per symbol."""
        return RdsDemod(self.lib)

    def rds_demod_taps(self):
        bp = np.empty(2 * 1024, np.float32)
        bank = np.empty(128 * 8, np.float32)
        n = self.lib.orc_rdsdemod_taps(bp.ctypes.data_as(C.c_void_p), 1024, bank.ctypes.data_as(C.c_void_p))
        return bp[: 2 * n].view(np.complex64).copy(), bank.reshape(128, 8)

    def rds_group_decode(self, bits):
        """reference build only: the radio module's RDS group decoder (rds.cpp) on a bit stream -> (PI code or None, PS name or None)"""
        L = self.lib
        L.orc_rdsdec_create.restype = C.c_void_p
        L.orc_rdsdec_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_rdsdec_pi.argtypes = [C.c_void_p]
        L.orc_rdsdec_ps.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.orc_rdsdec_free.argtypes = [C.c_void_p]
        b = np.ascontiguousarray(bits, np.uint8)
        h = L.orc_rdsdec_create()
        L.orc_rdsdec_process(h, b.ctypes.data_as(C.c_void_p), int(b.size))
        pi = L.orc_rdsdec_pi(h)
        buf = C.create_string_buffer(64)
        n = L.orc_rdsdec_ps(h, buf, 64)
        L.orc_rdsdec_free(h)
        return (pi if pi >= 0 else None), (buf.value.decode("latin-1") if n >= 0 else None)

    # ---- spectrum branch ----
    def fft_params(self, sr, size, rate):
        skip, nz = C.c_int(), C.c_int()
        self.lib.orc_fft_params(sr, size, rate, C.byref(skip), C.byref(nz))
        return skip.value, nz.value

    def window_buf(self, win, nz):
        out = np.empty(nz, np.float32)
        self.lib.orc_window_buf(win, nz, out.ctypes.data_as(C.c_void_p))
        return out

    def fft_frame(self, size, nz, win, iq):
        """iq: complex64[nz] -> float32[size] dB line (one-shot convenience)."""
        h = self.lib.orc_fft_create(size, nz, win)
        iq = np.ascontiguousarray(iq, np.complex64)
        out = np.empty(size, np.float32)
        self.lib.orc_fft_frame(h, iq.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        self.lib.orc_fft_free(h)
        return out

    def fft_raw(self, size, nz, win, iq):
        h = self.lib.orc_fft_create(size, nz, win)
        iq = np.ascontiguousarray(iq, np.complex64)
        out = np.empty(size, np.complex64)
        self.lib.orc_fft_raw(h, iq.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        self.lib.orc_fft_free(h)
        return out

    def zoom(self, offset, width, out_size, line):
        line = _f32(line)
        out = np.empty(out_size, np.float32)
        self.lib.orc_zoom(offset, width, line.size, out_size, line.ctypes.data_as(C.c_void_p),
                          out.ctypes.data_as(C.c_void_p))
        return out

    def hold(self, hold, latest, speed):
        hold = _f32(hold).copy()
        latest = _f32(latest)
        self.lib.orc_hold(hold.ctypes.data_as(C.c_void_p), latest.ctypes.data_as(C.c_void_p), hold.size, speed)
        return hold

    def pcm_compress(self, iq, pcm_type):
        """SampleStreamCompressor::process: complex64 -> packet bytes; pcm_type 0 int8, 1 int16, 2 float32"""
        x = np.ascontiguousarray(iq, np.complex64)
        out = np.zeros(8 + x.size * 8, np.uint8)
        n = self.lib.orc_pcm_compress(int(x.size), int(pcm_type), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out[:n].copy()

    def pcm_decompress(self, packet):
        b = np.ascontiguousarray(packet, np.uint8)
        out = np.zeros(max(b.size, 8), np.float32)
        n = self.lib.orc_pcm_decompress(int(b.size), b.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out[: 2 * n].view(np.complex64).copy()

    def export_convert(self, x, typ):
        a = np.ascontiguousarray(x, np.float32).reshape(-1)
        out = np.zeros(a.size, {0: np.uint8, 1: np.int16, 2: np.int32}[typ])
        self.lib.orc_export_convert(a.ctypes.data_as(C.c_void_p), int(a.size), int(typ), out.ctypes.data_as(C.c_void_p))
        return out

    def i16_to_f32(self, x):
        x = np.ascontiguousarray(x, np.int16)
        out = np.empty(x.size, np.float32)
        self.lib.orc_i16_to_f32(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), x.size)
        return out
