"""B200_TRACE timeline of a few device-resident steps of the bench workload (run with B200_TRACE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdrplusplus_b200 as sb
from sdrplusplus_b200 import lib
import bench
torch.cuda.set_device(0)
L = lib.load(); lib.check(L.b200_init(0))
opts = dict(a.split("=") for a in sys.argv[1:])
chunk = int(opts.pop("chunk", 1 << 24))
nofft = int(opts.pop("nofft", 0))
host = int(opts.pop("host", 0))
nsteps = int(opts.pop("steps", 8))
fe = sb.FrontEnd(bench.FS, chunk)
for k, v in opts.items():
    fe.set_option(k, int(v))
if not nofft:
    fe.set_fft(bench.FFT_SIZE, bench.FFT_RATE, 2)
ids = [fe.add_vfo(sb.VfoConfig.wfm(o)) for o in bench.OFFSETS]
if host:
    print("numa:", bench.bind_to_gpu_numa(0), file=sys.stderr)
    ins = [torch.zeros(2 * chunk, dtype=torch.int16, pin_memory=True) for _ in range(2)]
else:
    ins = [torch.rand(2 * chunk, device="cuda") * 2 - 1 for _ in range(3)]
outs = []
for _ in range(2):
    o = lib.Outputs(); keep = []
    for v in ids:
        cap = fe.vfo_max_out(v, chunk)
        t = torch.empty(2 * cap, dtype=torch.float32, pin_memory=True) if host else torch.empty(2 * cap, device="cuda")
        keep.append(t)
        o.vfo_out[v] = t.data_ptr(); o.vfo_cap[v] = cap
    nl = max(1, fe.fft_max_lines(chunk))
    t = torch.empty(nl * bench.FFT_SIZE, dtype=torch.float32, pin_memory=True) if host else torch.empty(nl * bench.FFT_SIZE, device="cuda")
    keep.append(t)
    o.fft_out = t.data_ptr(); o.fft_cap_lines = nl; o.out_mem = lib.MEM_HOST if host else lib.MEM_DEVICE
    outs.append((o, keep))
infl = 0
for i in range(nsteps):
    if host:
        fe.submit_ptr(ins[i % 2].data_ptr(), chunk, lib.FMT_CS16, lib.MEM_HOST, outs[i % 2][0])
    else:
        fe.submit_ptr(ins[i % 3].data_ptr(), chunk, lib.FMT_CF32, lib.MEM_DEVICE, outs[i % 2][0])
    infl += 1
    if infl == 2:
        fe.wait(); infl -= 1
while infl:
    fe.wait(); infl -= 1
fe.close()
