"""2 GPUs, one process each (torchrun): where does the N > 1 end-to-end arm lose its speed?  Every rank renders its round-robin
shard of one frame with host output at row-major positions into (a) its OWN page-locked full-frame buffer, (b) its own POSIX-shm
registered buffer, (c) ONE shm frame shared by both ranks."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import vkhrt_b200 as V
from vkhrt_b200.multi import SharedHostFrame
from multiprocessing import shared_memory

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
W, H = 2720, 1530
pos, idx = V.generate_groom(100000, 32, V.GROOM_CURLY)
vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
sc = V.Scene(pos, idx, technique=V.PHANTOM, device=lr).build()
n = W * H * 32

def run(name, ptr, sync_each=True, stride=world, first=rank):
    f = V.make_frame(vi, pi, W, H, tile_size=64, tile_first=first, tile_stride=stride, row_major_output=1 if stride > 1 else 0, output_memory=V.MEM_HOST)
    for _ in range(3):
        sc.render_into(f, ptr, None)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(20):
        sc.render_into(f, ptr, None)
        if sync_each:
            dist.barrier()
    torch.cuda.synchronize(); dist.barrier()
    ms = (time.time() - t0) / 20 * 1e3
    k = sc.timing()["trace_ms"]
    print(f"rank {rank} {name:52s} {ms:7.3f} ms/frame kernel {k:.3f} ms", flush=True)
    dist.barrier()

own = torch.empty(n, dtype=torch.uint8).pin_memory()
run("own cudaHostAlloc frame, barrier per frame", own.data_ptr())
run("own cudaHostAlloc frame, no barrier", own.data_ptr(), sync_each=False)
shm = shared_memory.SharedMemory(create=True, size=n); a = np.frombuffer(shm.buf, dtype=np.uint8); a[:] = 0
torch.cuda.cudart().cudaHostRegister(a.ctypes.data, n, 3)
run("own shm + cudaHostRegister frame, no barrier", a.ctypes.data, sync_each=False)
shared = SharedHostFrame(W * H)
shared.array[:] = 0
run("ONE shared shm frame, no barrier", shared.ptr, sync_each=False)
run("ONE shared shm frame, barrier per frame", shared.ptr)
# same shard on both ranks into the shared frame (both write the same tiles): is it the interleaving?
run("ONE shared frame, both ranks write shard 0", shared.ptr, sync_each=False, first=0)
if rank == 0:
    run_alone = True
dist.barrier()
dist.destroy_process_group()
os._exit(0)
