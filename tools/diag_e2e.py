"""GPU-box diagnostic: where does the end-to-end (host buffers) leg lose time against the PCIe ceiling?"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdrplusplus_b200 as sb
from sdrplusplus_b200 import lib
import bench

def ev():
    return torch.cuda.Event(enable_timing=True)

def main():
    torch.cuda.set_device(0)
    print("numa:", bench.bind_to_gpu_numa(0))
    L = lib.load(); lib.check(L.b200_init(0))
    nb = 64 << 20
    h_in = [torch.empty(nb, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    d_in = [torch.empty(nb, dtype=torch.uint8, device="cuda") for _ in range(2)]
    h_out = torch.empty(16 << 20, dtype=torch.uint8, pin_memory=True)
    d_out = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(label, with_d2h, with_kernel, reps=10):
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        with torch.cuda.stream(s1):
            e0.record(s1)
            for i in range(reps):
                d_in[i % 2].copy_(h_in[i % 2], non_blocking=True)
            e1.record(s1)
        with torch.cuda.stream(s2):
            for i in range(reps):
                if with_kernel:
                    for _ in range(8):
                        d_out.add_(1)
                if with_d2h:
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%-40s H2D %.1f GB/s (%.3f ms per 64 MiB)" % (label, reps * nb / ms / 1e6, ms / reps))
    run("H2D alone", False, False)
    run("H2D + concurrent D2H 16 MiB/step", True, False)
    run("H2D + concurrent kernels", False, True)
    run("H2D + kernels + D2H", True, True)

    # library e2e variants
    chunk = 1 << 24
    def lib_e2e(label, fft, nvfo, fmt, steps=10):
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            fe = sb.FrontEnd(bench.FS, chunk)
            fe.set_stream(stream.cuda_stream)
            if fft:
                fe.set_fft(bench.FFT_SIZE, bench.FFT_RATE, 2)
            ids = [fe.add_vfo(sb.VfoConfig.wfm(o)) for o in bench.OFFSETS[:nvfo]]
            bps = 4 if fmt == lib.FMT_CS16 else 8
            hin = [torch.zeros(chunk * bps, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            outs = []
            for _ in range(2):
                o = lib.Outputs(); keep = []
                for v in ids:
                    cap = fe.vfo_max_out(v, chunk)
                    t = torch.empty(2 * cap, dtype=torch.float32, pin_memory=True); keep.append(t)
                    o.vfo_out[v] = t.data_ptr(); o.vfo_cap[v] = cap
                nl = fe.fft_max_lines(chunk)
                t = torch.empty(max(nl, 1) * bench.FFT_SIZE, dtype=torch.float32, pin_memory=True); keep.append(t)
                o.fft_out = t.data_ptr(); o.fft_cap_lines = nl; o.out_mem = lib.MEM_HOST
                outs.append((o, keep))
            def loop(n):
                infl = 0
                for i in range(n):
                    fe.submit_ptr(hin[i % 2].data_ptr(), chunk, fmt, lib.MEM_HOST, outs[i % 2][0]); infl += 1
                    if infl == 2:
                        fe.wait(); infl -= 1
                while infl:
                    fe.wait(); infl -= 1
            loop(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            loop(steps)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / steps * 1e3
            print("%-40s %.3f ms/step  %.2f GS/s  (H2D alone would be %.3f ms)" % (label, dt, chunk / dt / 1e6, chunk * bps / 55.5e6))
            fe.close()
    lib_e2e("lib cs16: fft+8vfo", True, 8, lib.FMT_CS16)
    lib_e2e("lib cs16: 8vfo only", False, 8, lib.FMT_CS16)
    lib_e2e("lib cs16: fft only", True, 0, lib.FMT_CS16)
    lib_e2e("lib cs16: nothing (copy only)", False, 0, lib.FMT_CS16)
    lib_e2e("lib cf32: fft+8vfo", True, 8, lib.FMT_CF32)
    lib_e2e("lib cf32: nothing (copy only)", False, 0, lib.FMT_CF32)

main()
