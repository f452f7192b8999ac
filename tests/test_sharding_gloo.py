"""N > 1 host logic on CPU: world_size-2 and -3 `gloo` process groups run the same shard -> all_gather -> untile
sequence bench.py runs over NCCL, with the oracle standing in for the GPU kernels (test infrastructure only).
The gathered, untiled frame must be bit-identical to the un-sharded frame."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, W, H, T, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import vkhrt_b200 as V
    from vkhrt_b200.multi import TileSharding
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pos, idx = V.generate_groom(400, 8, V.GROOM_CURLY)
        vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
        lay = TileSharding(W, H, world, T)
        orc = O.OracleScene(pos, idx, technique=1)
        fo = O.make_frame(vi, pi, W, H, shade_mode=0, miss_rgb=(0.1, 0.2, 0.3), **lay.frame_kwargs(rank))
        hits, rgba, _ = orc.render(fo, n_out=lay.shard_pixels)
        assert hits.shape[0] == lay.shard_pixels == V.frame_local_pixels(V.make_frame(vi, pi, W, H, **lay.frame_kwargs(rank)))
        th = torch.from_numpy(hits.view(np.uint8).reshape(-1, 32).copy())
        ti = torch.from_numpy(rgba.copy())
        gh = torch.empty((world * lay.shard_pixels, 32), dtype=torch.uint8)
        gi = torch.empty((world * lay.shard_pixels, 4), dtype=torch.uint8)
        dist.all_gather_into_tensor(gh, th)
        dist.all_gather_into_tensor(gi, ti)
        full_h = lay.untile_host(gh.numpy())
        full_i = lay.untile_host(gi.numpy())
        # weak-scaling bookkeeping as bench.py does it: rays of all ranks / max-over-ranks time
        tmax = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        assert float(tmax) == float(world)
        if rank == 0:
            ref_h, ref_i, _ = orc.render(O.make_frame(vi, pi, W, H, shade_mode=0, miss_rgb=(0.1, 0.2, 0.3)))
            ok = full_h.tobytes() == ref_h.tobytes() and np.array_equal(full_i, ref_i) and int((ref_h["flags"] & 1).sum()) > 50
            open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H,T", [(2, 200, 120, 32), (3, 136, 72, 16)])
def test_shard_gather_untile_over_gloo(tmp_path, world, W, H, T):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, W, H, T, str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / "ok").read() == "1"
