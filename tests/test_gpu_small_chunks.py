"""Small chunks (the reference's own sizes, core/src/dsp/stream.h:9): the launches behind stage 1 are recorded, and a chunk whose
launch list was seen before replays a captured CUDA graph; the audio of a small chunk is stored by the kernels straight into
pinned host buffers from b200_host_alloc.  Both must be invisible in the results: bit-identical to the plain launches / the
copy-engine path, and within 1e-5 of the oracle."""
import ctypes as C

import numpy as np
import pytest

from util import rel_rms, noise_iq, fm_carrier

pytestmark = pytest.mark.gpu
TOL = 1e-5
FS = 100e6
OFFS = [-35e6, -15e6, 5e6, 25e6]


@pytest.fixture(scope="module")
def sb():
    import sdrplusplus_b200 as m
    from sdrplusplus_b200 import lib
    L = lib.load()
    assert L.b200_device_count() > 0
    assert L.b200_init(0) == 0
    return m


def _signal(n):
    x = noise_iq(n, 77, 0.01).copy()
    for k, o in enumerate(OFFS):
        x += fm_carrier(n, FS, o, tones=((1000.0 * (k + 1), 0.5), (5000.0, 0.3)))
    return x


def _run(sb, x, chunk, opts):
    fe = sb.FrontEnd(FS, chunk)
    for k, v in opts.items():
        fe.set_option(k, v)
    fe.set_fft(65536, 200.0, 2)
    ids = [fe.add_vfo(sb.VfoConfig.wfm(o)) for o in OFFS]
    outs, lines = fe.process_chunks(x, chunk)
    st = {k: fe.stat(k) for k in ("graph_hits", "graph_misses", "graphs")}
    fe.close()
    return [outs[i] for i in ids], lines, st


@pytest.mark.parametrize("chunk,min_hits", [(80000, 10), (51200, 10), (33333, 0)])      # 33333: the phases never repeat in 40 chunks
def test_graph_replay_is_bit_identical_and_matches_oracle(sb, oracle, report, chunk, min_hits):
    n = 40 * chunk
    x = _signal(n)
    ya, la, sa = _run(sb, x, chunk, {"graph": 0})
    yb, lb, sb_ = _run(sb, x, chunk, {"graph": 1})
    assert sa["graph_hits"] == 0 and sa["graphs"] == 0
    # the decimation / resampler phases repeat after a few chunks: most chunks must have replayed a graph
    assert sb_["graph_hits"] >= min_hits, sb_
    for a, b in zip(ya, yb):
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(la.view(np.uint32), lb.view(np.uint32))
    # ... and against the oracle, chunked the same way
    errs = []
    for o, y in zip(OFFS, yb):
        v, d = oracle.rxvfo(FS, 250e3, 150e3, o), oracle.wfm(75e3, 250e3)
        r = np.concatenate([d.process(v.process(x[i:i + chunk].view(np.float32))).reshape(-1, 2) for i in range(0, n, chunk)])
        assert r.shape == y.shape
        errs.append(rel_rms(y[1000:], r[1000:]))
    report["small_chunk_graph_replay_%d" % chunk] = {"wfm_audio_rel_rms": errs, "graph_hits": sb_["graph_hits"], "graphs": sb_["graphs"]}
    assert max(errs) < TOL, errs


def test_audio_stored_straight_into_pinned_host_buffers(sb):
    """Same stream through pinned b200_host_alloc buffers with host_direct on (kernels store over PCIe) and off (copy engine)."""
    from sdrplusplus_b200 import lib
    L = lib.load()
    L.b200_host_alloc.restype = C.c_void_p
    chunk, nch = 64000, 24
    x = _signal(chunk * nch)
    res = {}
    for mode in (0, 1):
        fe = sb.FrontEnd(FS, chunk)
        fe.set_option("host_direct", mode)
        ids = [fe.add_vfo(sb.VfoConfig.wfm(o)) for o in OFFS]
        hin = L.b200_host_alloc(chunk * 8)
        o = lib.Outputs()
        caps = {}
        for v in ids:
            caps[v] = fe.vfo_max_out(v, chunk)
            o.vfo_out[v] = L.b200_host_alloc(8 * caps[v]); o.vfo_cap[v] = caps[v]
        o.out_mem = lib.MEM_HOST
        acc = {v: [] for v in ids}
        for c in range(nch):
            C.memmove(hin, x[c * chunk:(c + 1) * chunk].ctypes.data, chunk * 8)
            fe.submit_ptr(hin, chunk, lib.FMT_CF32, lib.MEM_HOST, o)
            fe.wait()
            for v in ids:
                a = np.ctypeslib.as_array((C.c_float * (2 * o.vfo_count[v])).from_address(o.vfo_out[v]))
                acc[v].append(a.copy())
        res[mode] = [np.concatenate(acc[v]) for v in ids]
        fe.close()
        L.b200_host_free(C.c_void_p(hin))
        for v in ids:
            L.b200_host_free(C.c_void_p(o.vfo_out[v]))
    for a, b in zip(res[0], res[1]):
        assert a.size > 0 and a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_four_chunks_in_flight_same_bits(sb):
    """option "inflight" 4: submit runs up to four chunks ahead of wait (each with its own pinned input and output buffers);
    the stream that comes out is the one the synchronous call produces, bit for bit."""
    from sdrplusplus_b200 import lib
    L = lib.load()
    L.b200_host_alloc.restype = C.c_void_p
    chunk, nch, depth = 50000, 30, 4
    x = _signal(chunk * nch)
    # reference run: one chunk at a time
    fe = sb.FrontEnd(FS, chunk)
    fe.set_fft(65536, 400.0, 2)
    ids = [fe.add_vfo(sb.VfoConfig.wfm(o)) for o in OFFS]
    ref, ref_lines = fe.process_chunks(x, chunk)
    fe.close()
    fe = sb.FrontEnd(FS, chunk)
    fe.set_option("inflight", depth)
    fe.set_fft(65536, 400.0, 2)
    ids2 = [fe.add_vfo(sb.VfoConfig.wfm(o)) for o in OFFS]
    assert ids2 == ids
    nl = max(1, fe.fft_max_lines(chunk))
    hin = [L.b200_host_alloc(chunk * 8) for _ in range(depth)]
    outs = []
    for _ in range(depth):
        o = lib.Outputs()
        for v in ids:
            cap = fe.vfo_max_out(v, chunk)
            o.vfo_out[v] = L.b200_host_alloc(8 * cap); o.vfo_cap[v] = cap
        o.fft_out = L.b200_host_alloc(4 * nl * 65536); o.fft_cap_lines = nl; o.out_mem = lib.MEM_HOST
        outs.append(o)
    acc, lines = {v: [] for v in ids}, []

    def collect(k):
        o = outs[k % depth]
        for v in ids:
            acc[v].append(np.ctypeslib.as_array((C.c_float * (2 * o.vfo_count[v])).from_address(o.vfo_out[v])).copy())
        if o.fft_lines:
            lines.append(np.ctypeslib.as_array((C.c_float * (o.fft_lines * 65536)).from_address(o.fft_out)).copy().reshape(-1, 65536))
    done = 0
    for c in range(nch):
        C.memmove(hin[c % depth], x[c * chunk:(c + 1) * chunk].ctypes.data, chunk * 8)
        fe.submit_ptr(hin[c % depth], chunk, lib.FMT_CF32, lib.MEM_HOST, outs[c % depth])
        if c - done + 1 == depth:
            fe.wait(); collect(done); done += 1
    while done < nch:
        fe.wait(); collect(done); done += 1
    with pytest.raises(lib.B200Error):
        fe.wait()                                                    # nothing left in flight
    fe.close()
    for v in ids:
        y = np.concatenate(acc[v]).reshape(-1, 2)
        assert y.shape == ref[v].shape and np.array_equal(y.view(np.uint32), ref[v].view(np.uint32))
    got = np.concatenate(lines) if lines else np.empty((0, 65536), np.float32)
    assert got.shape == ref_lines.shape and np.array_equal(got.view(np.uint32), ref_lines.view(np.uint32))
    for p in hin:
        L.b200_host_free(C.c_void_p(p))
    for o in outs:
        for v in ids:
            L.b200_host_free(C.c_void_p(o.vfo_out[v]))
        L.b200_host_free(C.c_void_p(o.fft_out))
