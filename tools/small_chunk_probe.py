#!/usr/bin/env python
"""Where does a small chunk's time go?  BASELINE config 2 (100 MS/s, 1M-point FFT @ 20 fps, 8 WFM VFOs) at the reference's own
chunk sizes: device time per launch group (CUDA events), host time of b200_fe_submit by section, launches per chunk.
    python tools/small_chunk_probe.py > gpurun_out/small_chunk.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch                                                   # noqa: E402
import sdrplusplus_b200 as sb                                  # noqa: E402
from sdrplusplus_b200 import lib                               # noqa: E402

FS, FFT_SIZE, FFT_RATE = 100e6, 1 << 20, 20.0
OFFS = [-35e6, -25e6, -15e6, -5e6, 5e6, 15e6, 25e6, 35e6]


def run(csz, nchunks, opts, timers, host=False, depth=2):
    import ctypes as C
    import numpy as np
    L = lib.load()
    L.b200_host_alloc.restype = C.c_void_p
    fe = sb.FrontEnd(FS, csz)
    for k, v in opts.items():
        fe.set_option(k, v)
    fe.set_option("inflight", depth)
    fe.set_fft(FFT_SIZE, FFT_RATE, lib.WIN_NUTTALL)
    ids = [fe.add_vfo(sb.VfoConfig.wfm(o)) for o in OFFS]
    outs, keep = [], []
    for _ in range(depth):
        oo = lib.Outputs()
        nl = max(1, fe.fft_max_lines(csz))
        if host:
            for v in ids:
                c = fe.vfo_max_out(v, csz)
                oo.vfo_out[v] = L.b200_host_alloc(8 * c); oo.vfo_cap[v] = c
            oo.fft_out = L.b200_host_alloc(nl * FFT_SIZE * 4); oo.fft_cap_lines = nl; oo.out_mem = lib.MEM_HOST
        else:
            for v in ids:
                c = fe.vfo_max_out(v, csz)
                t = torch.empty(2 * c, device="cuda", dtype=torch.float32)
                keep.append(t); oo.vfo_out[v] = t.data_ptr(); oo.vfo_cap[v] = c
            t = torch.empty(nl * FFT_SIZE, device="cuda", dtype=torch.float32)
            keep.append(t); oo.fft_out = t.data_ptr(); oo.fft_cap_lines = nl; oo.out_mem = lib.MEM_DEVICE
        outs.append(oo)
    if host:
        # int16 IQ (file_source's format) in pinned host memory, 64 MiB walked chunk by chunk
        nbytes = 1 << 26
        hp = L.b200_host_alloc(nbytes)
        a = np.ctypeslib.as_array((C.c_int16 * (nbytes // 2)).from_address(hp))
        a[:] = (np.random.default_rng(1).standard_normal(nbytes // 2) * 2000).astype(np.int16)
        nslots = (nbytes // 4) // csz
        base, step, fmt, mem = hp, csz * 4, lib.FMT_CS16, lib.MEM_HOST
    else:
        big = torch.randn(1 << 26, device="cuda", dtype=torch.float32) * 0.1          # 256 MiB of IQ: larger than L2
        nslots = (big.numel() // 2) // csz
        base, step, fmt, mem = big.data_ptr(), csz * 8, lib.FMT_CF32, lib.MEM_DEVICE

    def loop(n):
        infl, hs, hw = 0, 0.0, 0.0
        for i in range(n):
            t0 = time.perf_counter()
            fe.submit_ptr(base + (i % nslots) * step, csz, fmt, mem, outs[i % depth])
            t1 = time.perf_counter()
            hs += t1 - t0
            infl += 1
            if infl == depth:
                fe.wait(); infl -= 1
                hw += time.perf_counter() - t1
        while infl:
            fe.wait(); infl -= 1
        return hs, hw
    loop(30)
    torch.cuda.synchronize()
    fe.set_option("time_s1", 1 if timers else 0)
    keys = ("host_ns_plan", "host_ns_fft", "host_ns_run", "host_ns_join", "host_ns_stage1", "host_ns_tail")
    h0 = {k: fe.stat(k) for k in keys}
    l0 = fe.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); hs, hw = loop(nchunks); e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / nchunks
    r = {"chunk": csz, "opts": opts, "inflight": depth, "mem": "host int16" if host else "device cf32", "graph_hits": fe.stat("graph_hits"), "graphs": fe.stat("graphs"), "us_per_chunk": us, "GS_per_s": csz / us / 1e3, "launches_per_chunk": (fe.launch_count() - l0) / nchunks,
         "host_submit_us": hs * 1e6 / nchunks, "host_wait_us": hw * 1e6 / nchunks,
         "host_sections_us": {k[8:]: (fe.stat(k) - h0[k]) / 1e3 / nchunks for k in keys}}
    if timers:
        g = {}
        for gi, name in ((0, "stage1"), (1, "behind_stage1"), (2, "spectrum")):
            ms, n = fe.group_stats(gi)
            g[name] = {"us": ms * 1e3 / n if n else None, "n": n}
        r["device_groups"] = g
    fe.close()
    if host:
        L.b200_host_free(C.c_void_p(hp))
        for oo in outs:
            for v in ids:
                L.b200_host_free(C.c_void_p(oo.vfo_out[v]))
            L.b200_host_free(C.c_void_p(oo.fft_out))
    return r


def h2d_probe():
    """cudaMemcpyAsync of one chunk's worth of pinned int16 IQ, back to back: what the copy engine gives small transfers."""
    out = {}
    for nbytes in (2 << 20, 4 << 20, 8 << 20, 64 << 20):
        h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        d.copy_(h, non_blocking=True); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            d.copy_(h, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        out[str(nbytes)] = {"us_per_copy": e0.elapsed_time(e1) * 1e3 / 50, "GB_per_s": 50 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9}
    return out


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":           # one configuration, for ncu's launch list
        L = lib.load()
        assert L.b200_init(0) == 0
        print(json.dumps(run(int(sys.argv[2]), 100, {"graph": 0}, False, False)))
        return
    from bench import bind_to_gpu_numa
    numa = bind_to_gpu_numa(0)
    L = lib.load()
    assert L.b200_init(0) == 0
    res = [{"numa": numa, "h2d_probe": h2d_probe()}]
    print(json.dumps(res[0]), file=sys.stderr)
    for csz, n in ((500000, 800), (1000000, 600), (1 << 21, 300)):
        for host in (False, True):
            for opts, depth in (({"graph": 0, "host_direct": 0}, 2), ({}, 2), ({}, 3), ({}, 4)):
                for timers in (False,):
                    try:
                        res.append(dict(run(csz, n, opts, timers, host, depth), timers=timers))
                    except Exception as ex:      # noqa: BLE001
                        res.append({"chunk": csz, "opts": opts, "error": repr(ex)})
                    print(json.dumps(res[-1]), file=sys.stderr)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
