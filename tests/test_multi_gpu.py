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
    from vkhrt_b200.multi import ShardedRenderer, SharedHostFrame
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
                    if mode == "peer":
                        # a MOVING camera: frame B is rendered while frame A's tensors are still unread; with one shared
                        # buffer the other ranks' frame-B stores would tear frame A under the gathering rank's eyes
                        vi2, pi2 = V.camera_matrices(position=(1.5, 151.0, 19.0), yaw=-84.0, pitch=-3.0, aspect=float(np.float32(W) / np.float32(H)))
                        ref2_h, ref2_i, _ = sc.render(V.make_frame(vi2, pi2, W, H, spp=spp))
                        f2 = sr.make_frame(vi2, pi2, stream.cuda_stream)
                        for _ in range(3):
                            a_h, a_i = sr.render(f, stream.cuda_stream)
                            b_h, b_i = sr.render(f2, stream.cuda_stream)
                            torch.cuda.synchronize()
                            dist.barrier()
                            if rank == 0:
                                ok &= np.array_equal(a_h.cpu().numpy().reshape(-1), ref_h.view(np.uint8))
                                ok &= np.array_equal(b_h.cpu().numpy().reshape(-1), ref2_h.view(np.uint8))
                                if rgba:
                                    ok &= np.array_equal(a_i.cpu().numpy(), ref_i) and np.array_equal(b_i.cpu().numpy(), ref2_i)
                            dist.barrier()
                    sr.close()
                # the end-to-end path at N > 1: every rank's kernel stores its records into ONE page-locked host frame
                shf = SharedHostFrame(W * H)
                for big in (False, True):                   # small frame: lane-bound kernel, zero-copy stores; big: pool kernel, line-wise
                    w2, h2 = (W, H) if not big else (1600, 1000)
                    if big and tech != V.PHANTOM:
                        continue
                    if big:
                        shf.close(); shf = SharedHostFrame(w2 * h2)
                    v3, p3 = V.camera_matrices(aspect=float(np.float32(w2) / np.float32(h2)))
                    full_h, _, _ = sc.render(V.make_frame(v3, p3, w2, h2), rgba=False)
                    shf.array[:] = 0xAB
                    dist.barrier()
                    fh = V.make_frame(v3, p3, w2, h2, tile_size=64, tile_first=rank, tile_stride=world, row_major_output=1, output_memory=V.MEM_HOST)
                    sc.render_into(fh, shf.ptr, None)
                    dist.barrier()
                    if rank == 0:
                        ok &= shf.hits().tobytes() == full_h.tobytes()
                    dist.barrier()
                shf.close()
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


def test_render_multi_on_real_gpus(V):
    """vkhrt_render_multi from ONE process over 2+ GPUs: peer stores into scenes[0]'s GPU (pixels; records into pageable memory)
    and zero-copy stores into a page-locked caller buffer (records) must both equal vkhrt_render bit for bit."""
    if V.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import ctypes as C
    import torch
    n = min(V.device_count(), 4)
    pos, idx = V.generate_groom(3000, 16, V.GROOM_CURLY)
    for tech, (W, H), spp in ((V.PHANTOM, (520, 300), 1), (V.PHANTOM, (1600, 1000), 1), (V.LSS, (520, 300), 2), (V.DOTS, (333, 222), 1)):
        vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
        scenes = [V.Scene(pos, idx, technique=tech, device=d).build() for d in range(n)]
        try:
            h0, i0, _ = scenes[0].render(V.make_frame(vi, pi, W, H, spp=spp, miss_rgb=(0.1, 0.2, 0.3)))
            for _ in range(2):
                h, img = V.render_multi(scenes, V.make_frame(vi, pi, W, H, spp=spp, miss_rgb=(0.1, 0.2, 0.3)))          # pageable outputs
                assert h.tobytes() == h0.tobytes() and np.array_equal(img, i0), (tech, W, H)
            # page-locked record buffer: every GPU stores straight into it
            ph = torch.empty((W * H, 32), dtype=torch.uint8).pin_memory()
            pi8 = torch.empty((W * H, 4), dtype=torch.uint8).pin_memory()
            arr = (C.c_void_p * n)(*[sc._h for sc in scenes])
            for want_img in (False, True):
                ph.fill_(0xCD)
                f = V.make_frame(vi, pi, W, H, spp=spp, miss_rgb=(0.1, 0.2, 0.3))
                rc = V.lib().vkhrt_render_multi(arr, n, C.byref(f), ph.data_ptr(), pi8.data_ptr() if want_img else None)
                assert rc == 0, V.lib().vkhrt_last_error()
                assert ph.numpy().tobytes() == h0.tobytes(), (tech, W, H, want_img)
                if want_img:
                    assert np.array_equal(pi8.numpy(), i0)
        finally:
            for sc in scenes:
                sc.close()
