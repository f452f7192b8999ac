"""N GPUs, one process each (torchrun): what does the HOST side take when every GPU's COPY ENGINE delivers 66 MB at once?
(the SM-issued 128-byte writes of the zero-copy path saturate at ~93 GB/s aggregate on the 8-GPU box: profiles/r02_scaling.txt)
  a) contiguous cudaMemcpyAsync into each rank's own page-locked buffer
  b) cudaMemcpy2DAsync per 64x64 tile into ONE shared page-locked frame at row-major positions (2 KB rows), this rank's round-robin tiles"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from vkhrt_b200.multi import SharedHostFrame, TileSharding

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 2073600 * 32
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.time()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / reps
    t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
dt = timed(lambda: host.copy_(dev, non_blocking=True))
if rank == 0: print(f"N={world} contiguous D2H, own pinned buffer: {n / dt / 1e9:.1f} GB/s per rank, {world * n / dt / 1e9:.1f} GB/s aggregate", flush=True)
# b) strided per-tile copies into one shared frame
W = int(np.ceil(1920 * np.sqrt(world) / 8) * 8); H = int(round(W * 1080 / 1920)); T = TileSharding.balanced_tile(W, world, 64)
lay = TileSharding(W, H, world, T)
shared = SharedHostFrame(W * H)
shard = torch.empty(lay.shard_pixels * 32, dtype=torch.uint8, device="cuda")
rt = torch.cuda.cudart()
tiles = lay.tiles_of_rank(rank)
import ctypes
lib = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
lib.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
st = torch.cuda.current_stream().cuda_stream
def tile_copies():
    for k, t in enumerate(tiles):
        ty, tx = divmod(t, lay.tiles_x)
        w = min(T, W - tx * T); h = min(T, H - ty * T)
        lib.cudaMemcpy2DAsync(shared.ptr + ((ty * T) * W + tx * T) * 32, W * 32, shard.data_ptr() + k * T * T * 32, T * 32, w * 32, h, 2, st)
dt = timed(tile_copies, reps=10)
moved = sum(min(T, W - (t % lay.tiles_x) * T) * min(T, H - (t // lay.tiles_x) * T) for t in tiles) * 32
if rank == 0: print(f"N={world} per-tile 2-D D2H into ONE shared frame ({W}x{H}, {len(tiles)} tiles of {T} per rank): {dt * 1e3:.3f} ms per frame, {moved / dt / 1e9:.1f} GB/s per rank, {world * moved / dt / 1e9:.1f} GB/s aggregate", flush=True)
dist.barrier(); shared.close(); dist.destroy_process_group(); os._exit(0)
