"""Where the reference build exists (oracle/_ref/libsdrpp_ref.so: the reference's own dsp headers over the restated
leaf layer), the plain-C restatement must be BIT-IDENTICAL to it on fresh random inputs, block by block."""
import numpy as np
import pytest

from util import noise_iq, fm_carrier

FS = 2.4e6


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("seed", [1, 2])
def test_blocks_bit_identical(oracle, ref_oracle, seed):
    n = 120000
    x = (noise_iq(n, seed, 0.3) + fm_carrier(n, FS, 300e3)).astype(np.complex64).view(np.float32)
    R, S = ref_oracle, oracle
    mk = [
        (lambda o: o.xlator(-300e3, FS), 12000), (lambda o: o.decim(2), 7777), (lambda o: o.decim(16), 12000),
        (lambda o: o.decim(128), 12000), (lambda o: o.resamp(FS, 250e3), 12000), (lambda o: o.resamp(FS, 24e3), 12000),
        (lambda o: o.resamp(48e3, 250e3), 3000), (lambda o: o.rxvfo(FS, 250e3, 150e3, 300e3), 12000),
        (lambda o: o.rxvfo(FS, 250e3, 250e3, -123456.0), 11999), (lambda o: o.rxvfo(FS, 15e3, 10e3, 1e5), 12000),
        (lambda o: o.dcblock_c(50.0 / FS), 12000),
    ]
    for f, ch in mk:
        assert _same(f(R).process_chunks(x, ch), f(S).process_chunks(x, ch))
    y = R.rxvfo(FS, 250e3, 150e3, 300e3).process_chunks(x, 12000)
    # RDS side output of BroadcastFM (rdsOut): the restatement against the reference's mono branch, and the reference's stereo
    # branch against its mono branch (the same three blocks behind the same discriminator)
    rds_ref = R.wfm_rds(75e3, 250e3).process_chunks(y, 1250)
    assert abs(rds_ref.size // 2 - (y.size // 2) // 50) <= 1 and _same(rds_ref, S.wfm_rds(75e3, 250e3).process_chunks(y, 1250))
    import ctypes as C
    from oracle.oracle import Block
    R.lib.orc_wfm_rds_stereo_create.restype = C.c_void_p
    R.lib.orc_wfm_rds_stereo_create.argtypes = [C.c_double, C.c_double]
    assert _same(Block(R.lib, R.lib.orc_wfm_rds_stereo_create(75e3, 250e3), 2, 2).process_chunks(y, 1250), rds_ref)
    mk2 = [
        (lambda o: o.quad(75e3, 250e3), 1250), (lambda o: o.wfm(75e3, 250e3), 1250), (lambda o: o.wfm(75e3, 250e3, False, False), 999),
        (lambda o: o.nfm(250e3, 12500.0, True), 1250), (lambda o: o.am(1, 10e3, 2e-4, 2e-5, 4e-4, 250e3), 1250),
        (lambda o: o.am(0, 10e3, 2e-4, 2e-5, 4e-4, 250e3), 1250), (lambda o: o.ssb(0, 2800.0, 250e3, 2e-4, 2e-5), 1250),
        (lambda o: o.ssb(1, 2800.0, 250e3, 2e-4, 2e-5), 1250), (lambda o: o.ssb(2, 4600.0, 250e3, 2e-4, 2e-5), 1250),
        (lambda o: o.deemph(50e-6, 48e3), 480), (lambda o: o.resamp_stereo(250e3, 48e3), 1250),
        (lambda o: o.wfm(75e3, 250e3, True, True), 1250), (lambda o: o.wfm(75e3, 250e3, True, False), 777),      # stereo branch: pilot PLL
        (lambda o: o.squelch(-20.0), 1250), (lambda o: o.squelch(-60.0), 999),
        # IF chain of the radio module (radio_module.h:90-96): noise blanker, FM IF noise reduction at its four presets
        (lambda o: o.noise_blanker(500.0 / 24000.0, 2.0), 1250), (lambda o: o.noise_blanker(500.0 / 250e3, 1.2), 999),
        (lambda o: o.fm_if(32), 1250), (lambda o: o.fm_if(9), 999), (lambda o: o.fm_if(15), 1250), (lambda o: o.fm_if(31), 640),
    ]
    for f, ch in mk2:
        assert _same(f(R).process_chunks(y, ch), f(S).process_chunks(y, ch))


def test_design_and_spectrum_bit_identical(oracle, ref_oracle):
    R, S = ref_oracle, oracle
    for a in ((15000.0, 4000.0, 250000.0, False), (5000.0, 500.0, 15000.0, False), (3125.0, 312.5, 50000.0, True)):
        assert _same(R.lowpass(*a), S.lowpass(*a))
    assert _same(R.highpass(300.0, 100.0, 48000.0), S.highpass(300.0, 100.0, 48000.0))
    assert _same(R.bandpass_c(18750, 19250, 3000, 250000, True).view(np.float32), S.bandpass_c(18750, 19250, 3000, 250000, True).view(np.float32))
    for r in (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
        assert R.decim_plan(r) == S.decim_plan(r)
        for st in range(len(R.decim_plan(r))):
            assert _same(R.decim_taps(r, st), S.decim_taps(r, st))
    for rates in ((2.4e6, 250e3), (100e6, 250e3), (1.024e9, 50e3), (8e6, 24e3), (44100.0, 48000.0)):
        assert R.resamp_plan(*rates) == S.resamp_plan(*rates)
        assert _same(R.resamp_taps(*rates), S.resamp_taps(*rates))
    x = noise_iq(65536, 7, 1.0)
    for N, nz, w in ((1024, 1024, 0), (65536, 65536, 2), (4096, 3000, 1)):
        assert _same(R.fft_frame(N, nz, w, x[:nz]), S.fft_frame(N, nz, w, x[:nz]))
    line = R.fft_frame(65536, 65536, 2, x)
    for args in ((0, 65536, 1000), (1000, 60000, 1280), (60000, 8000, 777)):
        assert _same(R.zoom(args[0], args[1], args[2], line), S.zoom(args[0], args[1], args[2], line))


def test_packet_and_export_formats_bit_identical(oracle, ref_oracle):
    """compressed sample stream (sample_stream_compressor.h / _decompressor.h) and the recorder conversions: the
    restatement against the reference's own code, plus known answers."""
    x = (noise_iq(5000, 3, 0.4) + fm_carrier(5000, FS, 300e3)).astype(np.complex64)
    for t in (0, 1, 2):
        pr, ps = ref_oracle.pcm_compress(x, t), oracle.pcm_compress(x, t)
        assert np.array_equal(pr, ps)
        assert np.array_equal(ref_oracle.pcm_decompress(pr).view(np.uint32), oracle.pcm_decompress(ps).view(np.uint32))
        y = oracle.pcm_decompress(ps)
        # the scaler is the largest sample VALUE (not magnitude): samples inside (-max, max) come back within half a
        # step, more negative ones saturate -- the reference's behaviour, kept
        xf, yf = x.view(np.float32), y.view(np.float32)
        mx = float(np.max(xf))
        step = {0: mx / 128.0, 1: mx / 32768.0, 2: 0.0}[t]
        inside = np.abs(xf) < mx * (1.0 - 1.0 / 128.0)
        assert np.max(np.abs(yf[inside] - xf[inside])) <= step * 0.51 + 1e-9
        assert ps.size == 8 + x.size * {0: 2, 1: 4, 2: 8}[t]
        assert int(ps[2]) + 256 * int(ps[3]) == t and int(ps[0]) == 0 and int(ps[1]) == 0
    a = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 2.0, -2.0, 1.5e-5, 0.25000763], np.float32)
    for t in (0, 1, 2):
        assert np.array_equal(ref_oracle.export_convert(a[:5] if t == 0 else a, t), oracle.export_convert(a[:5] if t == 0 else a, t))
    assert list(oracle.export_convert(a, 1)[:7]) == [0, 16384, -16384, 32767, -32767, 32767, -32768]     # rintf: ties to even, saturation
    assert list(oracle.export_convert(a[:5], 0)) == [128, 191, 64, 255, 1]


def test_rds_demod_bit_identical(oracle, ref_oracle):
    """RDSDemod (decoder_modules/radio/src/rds_demod.h: FastAGC, two Costas loops, band-pass, M&M clock recovery, slicer,
    differential decoder): the restatement against the reference's own class, tap sets, soft values and bits, across chunk
    sizes and a reset; the loops lock and the bits come back."""
    from util import rds_baseband
    bp_r, bank_r = ref_oracle.rds_demod_taps()
    bp_s, bank_s = oracle.rds_demod_taps()
    assert bp_r.size == 190 and _same(bp_r.view(np.float32), bp_s.view(np.float32)) and _same(bank_r, bank_s)
    x, bits = rds_baseband(3000, 11)
    for chunk in (x.size, 839, 25, 1):
        xs = x if chunk > 1 else x[:1500]
        R, S = ref_oracle.rds_demod(), oracle.rds_demod()
        sr, hr = R.process_chunks(xs, chunk)
        ss, hs = S.process_chunks(xs, chunk)
        assert sr.size == ss.size and abs(sr.size - xs.size * 1187.5 / 5000.0) <= 4
        assert _same(sr, ss) and np.array_equal(hr, hs)
        if chunk == 839:
            R.reset(); S.reset()
            s2, h2 = S.process_chunks(xs, chunk)
            r2, g2 = R.process_chunks(xs, chunk)
            assert _same(s2, r2) and np.array_equal(h2, g2)
            # same start state after a reset, apart from the samples MM::reset leaves in its work buffer
            assert np.array_equal(h2[8:], hs[8:])
    # the decoded stream carries the transmitted bits (after the loops have settled), at some fixed delay
    S = oracle.rds_demod()
    _, hard = S.process_chunks(x, 1000)
    tail = hard[600:]
    best = max(np.mean(tail[: 2000] == bits[d: d + 2000]) for d in range(560, 640))
    assert best > 0.99, best


def test_rds_bits_are_what_the_reference_group_decoder_reads(oracle, ref_oracle):
    """An RDS subcarrier carrying four type-0A groups (PI 0xB200, PS name 'B200 DSP', checkwords per IEC 62106) on an FM
    carrier -> rdsOut branch -> RDSDemod (restatement, and the reference's own class) -> the reference's OWN group decoder
    (decoder_modules/radio/src/rds.cpp, compiled from where it lies) finds the PI code and the name."""
    from util import rds_mpx_iq, rds_group_bits
    bits = rds_group_bits(0xB200, "B200 DSP", 6)
    assert ref_oracle.rds_group_decode(bits) == (0xB200, "B200 DSP")               # the generator builds valid blocks
    x, _ = rds_mpx_iq(0, 3, bits=bits)
    for O in (oracle, ref_oracle):
        y = O.wfm_rds(75e3, 250e3).process_chunks(x.view(np.float32), 12500).view(np.complex64)
        _, hard = O.rds_demod().process_chunks(y, 250)
        assert ref_oracle.rds_group_decode(hard) == (0xB200, "B200 DSP")
