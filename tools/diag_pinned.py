"""Which NUMA node do torch pinned buffers land on, and how fast is H2D/D2H for each of them?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
torch.cuda.set_device(0)
print("numa:", bench.bind_to_gpu_numa(0))
libc = ctypes.CDLL(None, use_errno=True)
def node_of(addr):
    page = ctypes.c_void_p(addr & ~4095)
    status = ctypes.c_int(-99)
    rc = libc.syscall(279, 0, 1, ctypes.byref(page), None, ctypes.byref(status), 0)   # move_pages (x86-64)
    return status.value if rc == 0 else "err%d" % ctypes.get_errno()
bufs = [torch.zeros(64 << 20, dtype=torch.uint8, pin_memory=True) for _ in range(6)]
d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for i, b in enumerate(bufs):
    nodes = [node_of(b.data_ptr() + off) for off in (0, 16 << 20, 48 << 20, (64 << 20) - 4096)]
    for direction in ("h2d", "d2h"):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            if direction == "h2d": d.copy_(b, non_blocking=True)
            else: b.copy_(d, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        print("buf %d nodes %s %s %.1f GB/s" % (i, nodes, direction, 3 * (64 << 20) / e0.elapsed_time(e1) / 1e6))
os.system("cat /sys/devices/system/node/online; ls /sys/devices/system/node/ | head; nvidia-smi topo -m 2>/dev/null | head -20")
