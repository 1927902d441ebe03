"""D2H bandwidth into 1 GiB host buffers: torch pin_memory() vs an anonymous mmap with MADV_HUGEPAGE registered with
cudaHostRegister (fewer IOMMU / IOTLB entries per byte on a virtualised host)."""
import ctypes, mmap, os, sys
import torch

N = 1 << 30
dev = torch.device("cuda", 0)
src = torch.empty(N, dtype=torch.uint8, device=dev)
src.zero_()
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), flush=True)


def bw(dst, label):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(src, non_blocking=True); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{label}: {N / 1e6 / best:.1f} GB/s", flush=True)


a = torch.empty(N, dtype=torch.uint8).pin_memory()
bw(a, "torch pin_memory (cudaHostAlloc)")
del a

libc = ctypes.CDLL("libc.so.6", use_errno=True)
m = mmap.mmap(-1, N + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
buf = (ctypes.c_char * len(m)).from_buffer(m)
addr = ctypes.addressof(buf)
aligned = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
MADV_HUGEPAGE = 14
r = libc.madvise(ctypes.c_void_p(aligned), ctypes.c_size_t(N), MADV_HUGEPAGE)
print("madvise rc", r, flush=True)
ctypes.memset(aligned, 0, N)          # touch: fault the pages in (as huge pages if THP allows)
t = torch.frombuffer(buf, dtype=torch.uint8, count=N, offset=aligned - addr)
rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), N, 0)
print("cudaHostRegister rc", rc, "is_pinned", t.is_pinned(), flush=True)
bw(t, "mmap + MADV_HUGEPAGE + cudaHostRegister")
try:
    print("AnonHugePages:", [l for l in open("/proc/meminfo") if "AnonHugePages" in l][0].strip())
except Exception as e:
    print(e)
torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
