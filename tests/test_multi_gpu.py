"""Multi-GPU frame assembly on real GPUs (needs >= 2 devices; skipped otherwise): one process per GPU over NCCL,
'peer' mode (kernels store into rank 0's frame buffer over NVLink) and 'gather' mode (all_gather + untile) must both
reproduce the single-GPU frame bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import vkhrt_b200 as V
    from vkhrt_b200.multi import ShardedRenderer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    try:
        pos, idx = V.generate_groom(3000, 16, V.GROOM_CURLY)
        W, H = 520, 300                                     # partial tiles on both axes
        vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        for tech, spp, rgba in ((V.PHANTOM, 1, False), (V.LSS, 2, True)):
            with V.Scene(pos, idx, technique=tech, device=rank) as sc:
                sc.build()
                ref_h, ref_i, _ = sc.render(V.make_frame(vi, pi, W, H, spp=spp))
                for mode in ("peer", "gather"):
                    sr = ShardedRenderer(sc, W, H, tile=64, spp=spp, want_rgba=rgba, device=dev, mode=mode)
                    f = sr.make_frame(vi, pi, stream.cuda_stream)
                    for _ in range(2):
                        oh, oi = sr.render(f, stream.cuda_stream)
                    torch.cuda.synchronize()
                    dist.barrier()
                    if rank == 0:
                        ok &= sr.mode == mode                # no silent fallback
                        ok &= np.array_equal(oh.cpu().numpy().reshape(-1), ref_h.view(np.uint8))
                        if rgba:
                            ok &= np.array_equal(oi.cpu().numpy(), ref_i)
                    dist.barrier()
                    sr.close()
        if rank == 0:
            open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


def test_peer_and_gather_assembly_match_single_gpu(tmp_path, V):
    if V.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = min(V.device_count(), 4)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / "ok").read() == "1"
