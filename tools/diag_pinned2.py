import ctypes, mmap, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sdrplusplus_b200 import lib
torch.cuda.set_device(0)
print("numa:", bench.bind_to_gpu_numa(0))
os.system("cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag; grep -E 'AnonHuge|HugePages_|Hugepagesize' /proc/meminfo")
L = lib.load(); lib.check(L.b200_init(0))
cudart = ctypes.CDLL("libcudart.so.12") if False else None
d = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
def rate(ptr, n, h2d=True):
    import ctypes as C
    t = torch.empty(0)
    # use torch's cudaMemcpyAsync through a tensor view of raw memory is awkward: use cuda-python? fall back to ctypes on libcudart
    return None
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
def bw(hptr, n, h2d):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.current_stream().cuda_stream
    e0.record()
    for _ in range(3):
        if h2d: rt.cudaMemcpyAsync(d.data_ptr(), hptr, n, 1, s)
        else: rt.cudaMemcpyAsync(hptr, d.data_ptr(), n, 2, s)
    e1.record(); torch.cuda.synchronize()
    return 3 * n / e0.elapsed_time(e1) / 1e6
n = 64 << 20
print("--- cudaHostAlloc (b200_host_alloc)")
L.b200_host_alloc.restype = ctypes.c_void_p
ptrs = [L.b200_host_alloc(n) for _ in range(10)]
for i, p in enumerate(ptrs):
    ctypes.memset(p, 1, n)
    print("hostalloc %d h2d %.1f d2h %.1f" % (i, bw(p, n, True), bw(p, n, False)))
print("--- mmap + MADV_HUGEPAGE + cudaHostRegister")
libc = ctypes.CDLL(None, use_errno=True)
libc.mmap.restype = ctypes.c_void_p
libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
for i in range(10):
    raw = libc.mmap(None, n + (2 << 20), 3, 0x22, -1, 0)       # PROT_READ|WRITE, MAP_PRIVATE|MAP_ANONYMOUS
    p = (raw + (2 << 20) - 1) & ~((2 << 20) - 1)
    libc.madvise(ctypes.c_void_p(p), ctypes.c_size_t(n), 14)     # MADV_HUGEPAGE
    ctypes.memset(p, 1, n)
    rc = rt.cudaHostRegister(p, n, 0)
    print("hugereg %d rc=%d h2d %.1f d2h %.1f" % (i, rc, bw(p, n, True), bw(p, n, False)))
os.system("grep -E 'AnonHuge' /proc/meminfo")
