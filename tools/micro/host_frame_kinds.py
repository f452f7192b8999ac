"""Which kind of page-locked host memory can the traversal kernel deliver records into at full speed?  (round 2: the N > 1 e2e
arm first used POSIX shared memory + cudaHostRegister and ran at 11 GB/s per GPU.)  One GPU, C2, vkhrt_render with host output."""
import ctypes, mmap, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import vkhrt_b200 as V
from multiprocessing import shared_memory

W, H = 1920, 1080
pos, idx = V.generate_groom(100000, 32, V.GROOM_CURLY)
vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
sc = V.Scene(pos, idx, technique=V.PHANTOM).build()
n = W * H * 32
rt = torch.cuda.cudart()
torch.cuda.init()

def run(name, ptr):
    f = V.make_frame(vi, pi, W, H, output_memory=V.MEM_HOST)
    for _ in range(3):
        sc.render_into(f, ptr, None)
    t0 = time.time()
    for _ in range(20):
        sc.render_into(f, ptr, None)
    ms = (time.time() - t0) / 20 * 1e3
    print(f"{name:48s} {ms:7.3f} ms/frame  {W * H / ms / 1e3:8.1f} Mrays/s  kernel {sc.timing()['trace_ms']:.3f} ms", flush=True)

t = torch.empty(n, dtype=torch.uint8).pin_memory(); run("torch pin_memory (cudaHostAlloc)", t.data_ptr())
hb = V.api.HostBuffer(n, np.uint8); run("vkhrt_host_alloc (portable|mapped)", hb.ptr)
shm = shared_memory.SharedMemory(create=True, size=n)
a = np.frombuffer(shm.buf, dtype=np.uint8); a[:] = 0
print("register shm rc", rt.cudaHostRegister(a.ctypes.data, n, 3)); run("POSIX shm + cudaHostRegister", a.ctypes.data)
rt.cudaHostUnregister(a.ctypes.data)
m = mmap.mmap(-1, n + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
b = np.frombuffer(m, dtype=np.uint8)
base = (b.ctypes.data + (2 << 20) - 1) & ~((2 << 20) - 1)
try:
    m.madvise(mmap.MADV_HUGEPAGE)
except Exception as e:
    print("madvise", e)
b[:] = 0
print("register anon rc", rt.cudaHostRegister(base, n, 3)); run("anonymous mmap + MADV_HUGEPAGE + cudaHostRegister", base)
rt.cudaHostUnregister(base)
c = np.zeros(n + 4096, np.uint8); pc = (c.ctypes.data + 4095) & ~4095
print("register heap rc", rt.cudaHostRegister(pc, n, 3)); run("numpy heap + cudaHostRegister", pc)
for pth in ("/sys/kernel/mm/transparent_hugepage/enabled", "/sys/kernel/mm/transparent_hugepage/shmem_enabled"):
    try:
        print(pth, open(pth).read().strip())
    except Exception as e:
        print(pth, e)
