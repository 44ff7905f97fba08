"""Row (b) of SURVEY.md section 8, source level: the reference's OWN radio-module demodulator wrappers compile and link,
unchanged, against the adapter headers in sdrplusplus_b200/host/dsp (same class names, signatures and stream contract), and
on a GPU they produce the oracle's audio.

The wrapper headers are read where they lie under /root/reference through a symlink farm in build/ (nothing is copied into
the repository); `demod.h` there is the reference's file minus the CW demodulator (not on the path) and, for the six plain
wrappers, minus wfm.h -- which is built separately with the RDS chain behind it (build_radio_wfm).  The GPU box has no /root/reference: it runs the
binary built here (build/ travels with the snapshot like the library does)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RADIO = "/root/reference/decoder_modules/radio/src"
FARM = os.path.join(ROOT, "build", "radio_compile")
EXE = os.path.join(ROOT, "build", "radio_wrappers")
WRAPPERS = ["nfm.h", "am.h", "usb.h", "lsb.h", "dsb.h", "raw.h"]


def build_radio_wrappers():
    """symlink farm + g++; returns the executable path.  Raises on a compile error."""
    os.makedirs(os.path.join(FARM, "demodulators"), exist_ok=True)
    for h in WRAPPERS:
        dst = os.path.join(FARM, "demodulators", h)
        if os.path.islink(dst) or os.path.exists(dst):
            os.remove(dst)
        os.symlink(os.path.join(REF_RADIO, "demodulators", h), dst)
    with open(os.path.join(REF_RADIO, "demod.h")) as f:
        lines = f.read().splitlines()
    kept = [l for l in lines if not ("demodulators/wfm.h" in l or "demodulators/cw.h" in l)]
    assert len(kept) == len(lines) - 2
    with open(os.path.join(FARM, "demod.h"), "w") as f:         # generated, lives in build/ only
        f.write("\n".join(kept) + "\n")
    libdir = os.path.join(ROOT, "sdrplusplus_b200")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wno-unused-variable", "-Wno-unused-but-set-variable",
           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "sdrplusplus_b200", "host"), "-I", FARM,
           "-I", os.path.join(ROOT, "include"), "-o", EXE, os.path.join(ROOT, "tests", "stubs", "radio_wrappers.cpp"),
           "-L", libdir, "-lb200dsp", "-Wl,-rpath," + libdir, "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference radio wrappers do not compile against host/dsp:\n" + r.stdout)
    return EXE


EXE_WFM = os.path.join(ROOT, "build", "radio_wfm")
FARM_WFM = os.path.join(ROOT, "build", "radio_compile_wfm")


def build_radio_wfm():
    """The reference's WFM wrapper (demodulators/wfm.h: BroadcastFM + RDSDemod + Handler sinks + Reshaper + the RDS group
    decoder rds.cpp) against host/: a symlink farm of its own, `demod.h` = the reference's demod.h minus the CW demodulator, `rds_demod.h`
    = OUR adapter in the place of the module's own file, rds.h / rds.cpp the reference's own, read where they lie."""
    FARM = FARM_WFM                                             # its own farm: demod.h differs from the other build's
    os.makedirs(os.path.join(FARM, "demodulators"), exist_ok=True)
    links = {os.path.join("demodulators", h): os.path.join(REF_RADIO, "demodulators", h) for h in WRAPPERS + ["wfm.h"]}
    links["rds.h"] = os.path.join(REF_RADIO, "rds.h")
    links["rds_demod.h"] = os.path.join(ROOT, "sdrplusplus_b200", "host", "radio", "rds_demod.h")
    for rel, target in links.items():
        dst = os.path.join(FARM, rel)
        if os.path.islink(dst) or os.path.exists(dst):
            os.remove(dst)
        os.symlink(target, dst)
    with open(os.path.join(REF_RADIO, "demod.h")) as f:
        lines = f.read().splitlines()
    kept = [l for l in lines if "demodulators/cw.h" not in l]
    assert len(kept) == len(lines) - 1
    with open(os.path.join(FARM, "demod.h"), "w") as f:         # generated, lives in build/ only
        f.write("\n".join(kept) + "\n")
    libdir = os.path.join(ROOT, "sdrplusplus_b200")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-sign-compare",
           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "sdrplusplus_b200", "host"), "-I", FARM,
           "-I", os.path.join(ROOT, "include"), "-o", EXE_WFM, os.path.join(ROOT, "tests", "stubs", "radio_wfm.cpp"),
           os.path.join(REF_RADIO, "rds.cpp"),
           "-L", libdir, "-lb200dsp", "-Wl,-rpath," + libdir, "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("the reference's WFM wrapper does not compile against host/:\n" + r.stdout)
    return EXE_WFM


@pytest.mark.skipif(not os.path.isdir(REF_RADIO), reason="needs /root/reference (build container only)")
def test_reference_wfm_wrapper_with_rds_compiles_against_adapters():
    exe = build_radio_wfm()
    from sdrplusplus_b200 import lib
    if lib.load().b200_device_count() == 0:
        out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
        assert out.returncode == 0, out.stdout
        assert "no CUDA device" in out.stdout and "no CPU fallback" in out.stdout


@pytest.mark.skipif(not os.path.isdir(REF_RADIO), reason="needs /root/reference (build container only)")
def test_reference_radio_wrappers_compile_against_adapters():
    exe = build_radio_wrappers()
    from sdrplusplus_b200 import lib
    if lib.load().b200_device_count() == 0:
        out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
        assert out.returncode == 0, out.stdout
        assert "no CUDA device" in out.stdout and "no CPU fallback" in out.stdout


def test_adapter_headers_cover_the_hot_path_blocks():
    """every reference header INTEGRATION.md's table names has its adapter in the tree"""
    host = os.path.join(ROOT, "sdrplusplus_b200", "host", "dsp")
    for rel in ["stream.h", "block.h", "processor.h", "sink.h", "source.h", "operator.h", "hier_block.h", "chain.h", "types.h",
                "channel/rx_vfo.h", "channel/frequency_xlator.h", "multirate/rational_resampler.h", "multirate/power_decimator.h",
                "filter/fir.h", "filter/deephasis.h", "taps/tap.h", "taps/low_pass.h", "taps/high_pass.h",
                "demod/quadrature.h", "demod/fm.h", "demod/am.h", "demod/ssb.h", "demod/broadcast_fm.h",
                "convert/mono_to_stereo.h", "convert/complex_to_stereo.h", "compression/sample_stream_compressor.h", "b200/frontend.h",
                "sink/handler_sink.h", "buffer/reshaper.h", "noise_reduction/noise_blanker.h", "noise_reduction/fm_if.h", "noise_reduction/power_squelch.h"]:
        assert os.path.exists(os.path.join(host, rel)), rel


@pytest.mark.gpu
def test_reference_radio_wrappers_run_on_gpu(tmp_path, oracle, report):
    if not os.path.exists(EXE):
        pytest.skip("build/radio_wrappers not built (needs /root/reference at build time)")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import noise_iq, fm_carrier, am_carrier, ssb_tone
    n, chunk = 200000, 5000
    # one input file, read by every wrapper at its own IF rate (50 k / 15 k / 24 k / 48 k): a carrier near DC with
    # slow FM + AM so that each demodulator has something well-conditioned to work on
    t = np.arange(n, dtype=np.float64)
    x = (0.3 * (1.0 + 0.5 * np.sin(2 * np.pi * t / 97.0)) * np.exp(1j * (2 * np.pi * 0.02 * t + 1.5 * np.sin(2 * np.pi * t / 61.0)))).astype(np.complex64)
    x += noise_iq(n, 41, 0.002)
    path = os.path.join(str(tmp_path), "iq.f32")
    x.tofile(path)
    out = subprocess.run([EXE, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = {}
    for l in out.stdout.strip().splitlines():
        f = l.split()
        got[f[0]] = (int(f[1]), float(f[2]))
    xf = x.view(np.float32)
    def run(blk):
        return np.concatenate([blk.process(xf[2 * i: 2 * (i + chunk)]) for i in range(0, n, chunk)]).astype(np.float64)
    # the wrappers' own parameterisation (demodulators/nfm.h:29,56; am.h:34,76; usb.h:34,70): IF rates 50 k / 15 k / 24 k
    want = {
        "NFM": run(oracle.nfm(50000.0, 12500.0, True)),
        "AM": run(oracle.am(1, 10000.0, 50.0 / 15000.0, 5.0 / 15000.0, 100.0 / 15000.0, 15000.0)),
        "USB": run(oracle.ssb(0, 2800.0, 24000.0, 50.0 / 24000.0, 5.0 / 24000.0)),
        "LSB": run(oracle.ssb(1, 2800.0, 24000.0, 50.0 / 24000.0, 5.0 / 24000.0)),
        "DSB": run(oracle.ssb(2, 4600.0, 24000.0, 50.0 / 24000.0, 5.0 / 24000.0)),
        "RAW": xf.astype(np.float64),
    }
    errs = {}
    for k, y in want.items():
        assert k in got, (k, out.stdout)
        assert got[k][0] == n, (k, got[k])
        ref = float(np.sum(np.abs(y)))
        errs[k] = abs(got[k][1] - ref) / ref
    report["radio_wrappers_checksum_rel_err"] = errs
    for k, e in errs.items():
        assert e < 1e-5, (k, e)


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="five worker-thread blocks behind the wrapper; built late in the round, first GPU run is the driver's")
def test_reference_wfm_wrapper_decodes_rds_on_gpu(tmp_path, report):
    """FM carrier with an RDS subcarrier carrying PI 0xB200 / PS 'B200 DSP' -> the reference's WFM wrapper on the adapter
    classes (BroadcastFM + RDSDemod on the GPU) -> the reference's own group decoder -> the name the module would display."""
    if not os.path.exists(EXE_WFM):
        pytest.skip("build/radio_wfm not built (needs /root/reference at build time)")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import rds_mpx_iq, rds_group_bits
    x, _ = rds_mpx_iq(0, 3, bits=rds_group_bits(0xB200, "B200 DSP", 6))
    x = x[: (x.size // 12500) * 12500]
    path = os.path.join(str(tmp_path), "if.f32")
    x.tofile(path)
    out = subprocess.run([EXE_WFM, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = dict(l.split(" ", 1) for l in out.stdout.strip().splitlines())
    report["radio_wfm_wrapper"] = lines
    assert int(lines["WFM"].split()[0]) == x.size
    assert lines["RDS"].strip() == "RDS: B200 DSP", lines
