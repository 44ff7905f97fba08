"""The C++ adapter headers (sdrplusplus_b200/host/dsp: reference class names and signatures over the C ABI) used the
way a reference module uses the real blocks: worker-thread blocks joined by dsp::stream<T>.  The program is built by
__graft_entry__.build(); here it runs on the GPU and its audio is compared with the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import rel_rms, noise_iq, fm_carrier

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_adapter")


def test_adapter_headers_compile():
    """CPU-side: the adapter headers are self-contained C++17 and only need include/b200dsp.h."""
    host = os.path.join(ROOT, "sdrplusplus_b200", "host")
    headers = sorted(os.path.relpath(os.path.join(d, f), host) for d, _, fs in os.walk(os.path.join(host, "dsp")) for f in fs if f.endswith(".h"))
    assert len(headers) >= 33
    headers.append(os.path.join("radio", "rds_demod.h"))      # the radio module's RDSDemod (decoder_modules/radio/src/rds_demod.h)
    for h in headers:                                   # every header of the mirrored tree, each on its own
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", host, "-x", "c++", os.path.join(host, h)],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, (h, r.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["blocks", "fused"])
def test_adapter_graph_matches_oracle(oracle, tmp_path, mode):
    if not os.path.exists(EXE):
        pytest.skip("build/test_adapter missing: run python __graft_entry__.py")
    fs, n, chunk = 2.4e6, 240000, 12000
    x = noise_iq(n, 31, 0.02).copy() + fm_carrier(n, fs, 300e3)
    src, dst = tmp_path / "iq.f32", tmp_path / "audio.f32"
    x.tofile(src)
    args = [EXE, str(src), str(dst)] + (["fused"] if mode == "fused" else [])
    r = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    got = np.fromfile(dst, np.float32).reshape(-1, 2)
    v, d = oracle.rxvfo(fs, 250e3, 150e3, 300e3), oracle.wfm(75e3, 250e3)
    ref = np.concatenate([d.process(v.process(x.view(np.float32)[2 * i: 2 * (i + chunk)])) for i in range(0, n, chunk)]).reshape(-1, 2)
    assert got.shape == ref.shape
    assert rel_rms(got[4000:], ref[4000:]) < 1e-5
    if mode == "fused":
        assert int(r.stdout.split()[1]) == 2          # two 65536-pt lines completed in 240000 samples at 20 fps


@pytest.mark.gpu
def test_adapter_rds_wiring_decodes_bits(oracle, tmp_path):
    """The radio module's RDS wiring on the adapter classes (demodulators/wfm.h:78-81): BroadcastFM(rdsOut) -> RDSDemod as
    worker-thread blocks; the decoded bits are the oracle chain's and the transmitted ones."""
    from util import rds_mpx_iq
    if not os.path.exists(EXE):
        pytest.skip("build/test_adapter missing: run python __graft_entry__.py")
    x, bits = rds_mpx_iq(1500, 3)
    x = x[: (x.size // 12500) * 12500]
    src, dst = tmp_path / "if.f32", tmp_path / "bits.u8"
    x.tofile(src)
    r = subprocess.run([EXE, str(src), str(dst), "rds"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    got = np.fromfile(dst, np.uint8)
    nb, ns = (int(v) for v in r.stdout.split()[:2])
    assert nb == ns == got.size
    oracle.set_rotator_mode(1)
    try:
        y = oracle.wfm_rds(75e3, 250e3).process_chunks(x.view(np.float32), 12500).view(np.complex64)
    finally:
        oracle.set_rotator_mode(0)
    _, ho = oracle.rds_demod().process_chunks(y, 250)
    assert abs(got.size - ho.size) <= 1
    n = min(got.size, ho.size)
    assert np.count_nonzero(got[300:n] != ho[300:n]) <= 2          # a symbol on the threshold may differ (tests/test_gpu_rds.py)
    assert max(np.mean(got[300:1300] == bits[k: k + 1000]) for k in range(200, 400)) > 0.995


@pytest.mark.parametrize("keep,skip,total,chunk", [(8, -5, 100, 7), (4, 3, 200, 5), (4096, 39 - 4096, 9000, 250), (16, 0, 160, 16)])
def test_reshaper_and_handler_glue_blocks_on_cpu(keep, skip, total, chunk):
    """dsp::buffer::Reshaper + dsp::sink::Handler (host/dsp/buffer/reshaper.h, host/dsp/sink/handler_sink.h) as worker-thread
    blocks, against the framing rule of core/src/dsp/buffer/reshaper.h:100-126: blocks of `keep` samples; skip > 0 drops samples
    between blocks, skip < 0 repeats the last -skip samples of a block at the head of the next (zeros before the first)."""
    exe = os.path.join(ROOT, "build", "test_glue")
    if not os.path.exists(exe):
        pytest.skip("build/test_glue missing: run python __graft_entry__.py")
    r = subprocess.run([exe, str(keep), str(skip), str(total), str(chunk)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    got = [np.array(l.split(), np.float64) for l in r.stdout.strip().splitlines()]
    x = np.arange(1, total + 1, dtype=np.float64)
    carry = min(-skip, keep) if skip < 0 else 0
    fresh = keep - carry
    if fresh < 1:
        fresh, carry = 1, keep - 1
    want, pos, prev = [], 0, np.zeros(keep)
    while pos + fresh <= total:
        blk = np.concatenate([prev[fresh:], x[pos: pos + fresh]]) if carry else x[pos: pos + fresh]
        want.append(blk)
        prev = blk
        pos += fresh + max(skip, 0)
    assert len(got) == len(want), (len(got), len(want))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
