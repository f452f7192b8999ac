#!/usr/bin/env python
"""bench.py — Mrays/s of the primary-ray hair-hit path (ray-gen + BVH traversal + intersection +
hit-record write) on B200, with the CPU oracle timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Main line — workload "c2" (BASELINE.json configs[1]): synthetic curly groom 100k strands x 32 segments
(3.2 M Phantom curves), 1920x1080 primary rays per GPU, Phantom intersector, hit buffer only.  N > 1 is weak
scaling: the BVH is replicated, the frame grows to N x 1080p pixels at fixed aspect and camera, 64x64 tiles are
dealt round-robin to ranks and every rank's kernel stores its records straight into rank 0's frame buffer over
NVLink.  One step = one frame.  `value` is timed with CUDA events on the launching stream with outputs resident
in HBM; `e2e` goes through vkhrt_render with HOST buffers (camera in, hit records out to page-locked memory; at
N > 1 all ranks deliver into ONE shared page-locked frame).  `parity` compares the frame the timed code produced
with the CPU oracle (outside the timed region); `strong_c5` adds BASELINE configs[4] — 64 M curves, a FIXED
3840x2160 x 64 spp frame sharded over the N ranks — to every line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (strands, segments, style, width, height, technique, spp, rgba)
    "c1": (10000, 16, "straight", 512, 512, "phantom", 1, False),
    "c2": (100000, 32, "curly", 1920, 1080, "phantom", 1, False),
    "c3": (100000, 32, "curly", 1920, 1080, "lss", 8, True),
    "c4": (1000000, 16, "curly", 1920, 1080, "dots", 1, False),
    "c5": (1000000, 64, "curly", 3840, 2160, "phantom", 64, True),
    # not a BASELINE config: C2's groom at 4K, to separate kernel throughput from launch ramp/tail effects
    "c2_4k": (100000, 32, "curly", 3840, 2160, "phantom", 1, False),
}
# P in B_ray = 64 N_int + P N_prim + W.  DOTS: a leaf is one 64-byte strip record holding the 4 triangles of
# a segment, so P = 16 bytes per triangle of a fetched strip (a triangle-per-leaf tree would read 36 B each)
PRIM_BYTES = {"phantom": 48, "lss": 32, "dots": 16}
LEAF_RECORD_BYTES = {"phantom": 64, "lss": 32, "dots": 64}
TECH_ID = {"phantom": 0, "lss": 1, "dots": 2}
N_PARITY = 65536


def frame_size(base_w, base_h, n):
    if n == 1:
        return base_w, base_h
    w = int(math.ceil(base_w * math.sqrt(n) / 8.0) * 8)
    h = int(round(w * base_h / base_w))
    return w, h


def workload_desc(name, n, w, h, tile=64):
    s, g, style, _, _, tech, spp, rgba = WORKLOADS[name]
    return (f"{name}: synthetic {style} groom {s} strands x {g} segments ({s * g} segs), {tech} intersector, "
            f"{w}x{h} primary rays x {spp} spp, {'hit buffer + RGBA8' if rgba else 'hit buffer only'}"
            + (f", {tile}x{tile} tiles round-robin over {n} GPUs" if n > 1 else ""))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_counters(name, tech):
    """ncu counters of the dominant kernel for this workload, from the committed summaries under profiles/ (round 2)."""
    p = os.path.join(ROOT, "profiles", "counters.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(f"{name}_{tech}")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            try:
                pw.append(float(c[3]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ask the affinity mask instead)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------------------------
# the checker legs: ONLY these functions touch oracle/ (CPU baseline, --impl reference, parity blocks).  They use the
# oracle's own groom and camera generators, so the reference arm never maps the product library.
# ------------------------------------------------------------------------------------------------------------------
def oracle_scene(name):
    from oracle import oracle as O
    s, g, style, w, h, tech, spp, rgba = WORKLOADS[name]
    pos, idx = O.generate_groom(s, g, O.GROOM_STRAIGHT if style == "straight" else O.GROOM_CURLY)
    t0 = time.time()
    orc = O.OracleScene(pos, idx, technique=TECH_ID[tech])
    return orc, time.time() - t0


def oracle_frame(name, w, h, spp=1):
    from oracle import oracle as O
    vi, pi = O.camera_matrices(aspect=float(np.float32(w) / np.float32(h)))
    return O.make_frame(vi, pi, w, h, spp=spp)


def time_oracle(orc, frame, sub, reps):
    times, stats = [], None
    for _ in range(reps):
        t0 = time.time()
        _, _, stats = orc.render(frame, hits=True, rgba=False, pixel_subset=sub, stats=True, n_threads=host_threads())
        times.append(time.time() - t0)
    return times, stats


def parity_block(orc, name, w, h, spp, want_rgba, sub, hits_gpu_sub, rgba_gpu_sub):
    """Oracle over the pixel subset `sub` of the frame the GPU produced -> the `parity` object of the line."""
    from oracle.parity import parity_metrics
    ho, io, _ = orc.render(oracle_frame(name, w, h, spp), hits=True, rgba=want_rgba, pixel_subset=sub, n_threads=host_threads())
    m = parity_metrics(hits_gpu_sub, ho, rgba_gpu_sub if want_rgba else None, io if want_rgba else None)
    m["against"] = "CPU oracle (oracle/vkhrt_oracle.cpp), its own LBVH over its own copy of the groom"
    m["pixels"] = f"{len(sub)} seeded stratified pixels of the {w}x{h} frame, {spp} spp"
    return m


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  vkhrt has none (Vulkan RT, Windows only)
    and cannot be compiled here, so this is the oracle port on all host threads (oracle/vkhrt_oracle.cpp)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.parity import stratified_pixels
    name = args.workload
    s, g, style, w, h, tech, spp, rgba = WORKLOADS[name]
    orc, build_s = oracle_scene(name)
    frame = oracle_frame(name, w, h)
    n_sample = w * h if name == "c1" else 262144           # C1 is the full-frame CPU run
    sub = stratified_pixels(w, h, n_sample)
    time_oracle(orc, frame, sub[:4096], 1)
    for _ in range(args.warmup):
        time_oracle(orc, frame, sub, 1)
    t0 = time.time()
    times, stats = time_oracle(orc, frame, sub, args.steps)
    total = time.time() - t0
    mrays = len(sub) * args.steps / total / 1e6
    cores = host_threads()
    sample = f"{len(sub)} stratified pixels of the {w}x{h} frame per step (1 spp)"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s primary-ray hair hits", "value": mrays, "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(name, 1, w, h), "sample": sample, "oracle_build_s": build_s},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = mmzala/vkhrt has no CPU path and cannot be built here (Vulkan RT pipeline, Windows); "
                "this arm times the line-for-line C++ port of its shaders (oracle/), OpenMP over all host threads; "
                "grooms and cameras come from the oracle's own generators (the product library is not loaded)",
    }))


# ------------------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: 64 M curves, FIXED 3840x2160 x 64 spp frame sharded over the N ranks (strong scaling)
# ------------------------------------------------------------------------------------------------------------------
def strong_c5(V, torch, dist, world, rank, local_rank, dev, stream, steps, warmup):
    from vkhrt_b200.multi import ShardedRenderer
    s, g, style, W, H, tech, spp, want_rgba = WORKLOADS["c5"]
    pos, idx = V.generate_groom(s, g, V.GROOM_CURLY)
    vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
    scene = V.Scene(pos, idx, technique=V.PHANTOM, device=local_rank)
    del pos, idx
    scene.build()
    build_ms = scene.timing()["build_total_ms"]
    from vkhrt_b200.multi import TileSharding
    T = TileSharding.balanced_tile(W, world, 64)       # 3840 / 64 = 60 tiles per row would give 2 / 4 ranks vertical stripes
    sharded = ShardedRenderer(scene, W, H, tile=T, spp=spp, want_rgba=want_rgba, device=dev, mode="peer")
    fd = sharded.make_frame(vi, pi, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        sharded.render(fd, stream.cuda_stream)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        a.record(stream)
        o_hits, o_rgba = sharded.render(fd, stream.cuda_stream)
        b.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    frame_ms = [a.elapsed_time(b) for a, b in evs]
    t = torch.tensor([sum(frame_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    rays = W * H * spp
    value = rays * steps / (total_ms * 1e-3) / 1e6
    # per-rank time of its own shard (sample-0 traversal kernel x spp is not separable per sample; use the frame's device time
    # up to the completion signal): min / max over ranks = tile load imbalance
    mine = torch.tensor([sum(frame_ms) / steps], dtype=torch.float64, device=dev)
    lo, hi = mine.clone(), mine.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    # own-shard work only (no waiting for the others): one untimed extra frame without the completion all_reduce
    own = scene_shard_ms(V, torch, scene, sharded, fd, stream)
    own_t = torch.tensor([own], dtype=torch.float64, device=dev)
    own_lo, own_hi = own_t.clone(), own_t.clone()
    if world > 1:
        dist.all_reduce(own_lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(own_hi, op=dist.ReduceOp.MAX)

    # end to end: the assembled frame (records + pixels) arrives in rank 0's page-locked host memory
    e2e_steps = max(2, min(steps, 3))
    if rank == 0:
        h_hits = torch.empty((W * H, 32), dtype=torch.uint8).pin_memory()
        h_rgba = torch.empty((W * H, 4), dtype=torch.uint8).pin_memory()
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        o_hits, o_rgba = sharded.render(fd, stream.cuda_stream)
        if rank == 0:
            h_hits.copy_(o_hits, non_blocking=True)
            h_rgba.copy_(o_rgba, non_blocking=True)
        barrier()
    e2e_s = time.time() - t0
    e2e_value = rays * e2e_steps / e2e_s / 1e6

    out = None
    assembled = None
    if world > 1:
        # the assembled frame against rank 0's OWN single-GPU render of the whole frame (outside any timing)
        if rank == 0:
            whole = V.make_frame(vi, pi, W, H, spp=spp, output_memory=V.MEM_DEVICE, stream=stream.cuda_stream)
            d_h = torch.empty((W * H, 32), dtype=torch.uint8, device=dev)
            d_i = torch.empty((W * H, 4), dtype=torch.uint8, device=dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            scene.render_into(whole, d_h.data_ptr(), d_i.data_ptr())          # warm
            a.record(stream)
            scene.render_into(whole, d_h.data_ptr(), d_i.data_ptr())
            b.record(stream)
            torch.cuda.synchronize()
            single_ms = a.elapsed_time(b)
            tiles = ((H + T - 1) // T) * ((W + T - 1) // T)
            same_h = bool(torch.equal(d_h, o_hits)); same_i = bool(torch.equal(d_i, o_rgba))
            same_host = bool(torch.equal(h_hits, d_h.cpu())) and bool(torch.equal(h_rgba, d_i.cpu()))
            assembled = {"tiles_checked": tiles, "of_tiles": tiles, "hits_identical": same_h, "rgba_identical": same_i,
                         "host_frame_identical": same_host,
                         "against": "rank 0's own single-GPU render of the whole frame (every tile, every byte)",
                         "single_gpu_ms": single_ms, "single_gpu_mrays": rays / (single_ms * 1e-3) / 1e6}
        dist.barrier()
    if rank == 0:
        out = {"workload": workload_desc("c5", world, W, H, T), "scaling": "strong", "tile": T, "value": value, "unit": "Mrays/s",
               "ms_per_frame": total_ms / steps, "frames_timed": steps, "rays_per_frame": rays, "build_ms": build_ms,
               "frame_ms_min_max_over_ranks": [float(lo.item()), float(hi.item())],
               "own_shard_ms_min_max_over_ranks": [float(own_lo.item()), float(own_hi.item())],
               "tile_load_imbalance": float(own_hi.item()) / max(float(own_lo.item()), 1e-9),
               "e2e": {"value": e2e_value, "unit": "Mrays/s", "frames": e2e_steps, "d2h_bytes_per_frame": W * H * 36,
                       "note": "frame assembled in rank 0's HBM by peer stores, then ONE device->host copy of records + pixels into page-locked memory"},
               "parity_assembled": assembled, "clocks": clocks}
        if assembled:
            out["efficiency_vs_single_gpu_same_run"] = value / (world * assembled["single_gpu_mrays"])
    sharded.close()
    scene.close()
    return out


def scene_shard_ms(V, torch, scene, sharded, fd, stream):
    """device time of this rank's own shard (no completion signal): CUDA events around one render_into"""
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    scene.render_into(fd, sharded.hits_ptr, sharded.rgba_ptr)
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline duration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the produced frame (experiment sweeps)")
    ap.add_argument("--no-strong-c5", action="store_true", help="skip the BASELINE configs[4] block (experiment sweeps)")
    ap.add_argument("--c5-frames", type=int, default=4)
    ap.add_argument("--gather", default="peer", choices=["peer", "gather"],
                    help="N > 1: 'peer' = kernels store straight into the gathering rank's frame buffer over NVLink; 'gather' = NCCL all_gather + untile")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import vkhrt_b200 as V

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
    if V.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: vkhrt_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")       # keeps NCCL's version banner off stdout: rank 0 prints ONE line
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    s, g, style, bw, bh, tech, spp, want_rgba = WORKLOADS[name]
    tech_id = {"phantom": V.PHANTOM, "lss": V.LSS, "dots": V.DOTS}[tech]
    W, H = frame_size(bw, bh, world)
    pos, idx = V.generate_groom(s, g, V.GROOM_STRAIGHT if style == "straight" else V.GROOM_CURLY)
    vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
    scene = V.Scene(pos, idx, technique=tech_id, device=local_rank)
    scene.build()
    build_timing = scene.timing()
    n_leaves = scene.n_leaves

    # a non-default stream: the ABI treats a NULL stream handle as "use the scene's own stream"
    from vkhrt_b200.multi import ShardedRenderer, SharedHostFrame, TileSharding
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    T = TileSharding.balanced_tile(W, world, 64)
    sharded = ShardedRenderer(scene, W, H, tile=T, spp=spp, want_rgba=want_rgba, device=dev, mode=args.gather)
    fd = sharded.make_frame(vi, pi, stream.cuda_stream)
    n_local = sharded.layout.shard_pixels
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_device():
        # trace this rank's tiles; N > 1: the kernels' stores assemble the frame in rank 0's HBM (vkhrt_b200/multi.py)
        return sharded.render(fd, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel-level arm: outputs stay in HBM ----------------
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("VKHRT_BENCH_NO_SAMPLER")) else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = V.launch_count()
    barrier()
    t_wall0 = time.time()
    for a, b in evs:
        flush.zero_()                 # evict the scene from L2 between timed frames (not inside the events)
        a.record(stream)
        o_hits, o_rgba = step_device()
        b.record(stream)
    barrier()
    t_wall = time.time() - t_wall0
    launches = V.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    rays_per_step = W * H * spp
    value = rays_per_step * args.steps / (total_ms * 1e-3) / 1e6
    device_timing = scene.timing()          # per-stage CUDA events of the last device-resident frame
    # the frame the timed code produced (rank 0 holds it assembled): kept for the parity blocks below
    frame_hits = o_hits.cpu().numpy().reshape(-1).view(V.HIT_DTYPE) if rank == 0 else None
    frame_rgba = o_rgba.cpu().numpy() if (rank == 0 and want_rgba) else None

    # ---------------- end-to-end arm: the public call with HOST buffers ----------------
    # N = 1: vkhrt_render into a page-locked buffer.  N > 1: ONE page-locked frame shared by all ranks (POSIX shared memory
    # registered with CUDA in every rank); every rank's kernel stores its records at their row-major position over its own PCIe link.
    shared_host = None
    if world > 1 and not want_rgba:
        shared_host = SharedHostFrame(W * H)
        fh = V.make_frame(vi, pi, W, H, spp=spp, tile_size=T, tile_first=rank, tile_stride=world, row_major_output=1, output_memory=V.MEM_HOST)
        h_hits_ptr, h_rgba_ptr = shared_host.ptr, None
        e2e_bytes = n_local * 32
    else:
        fh = V.make_frame(vi, pi, W, H, spp=spp, tile_size=T, tile_first=rank, tile_stride=world, output_memory=V.MEM_HOST)
        h_hits = torch.empty((n_local, 32), dtype=torch.uint8).pin_memory()
        h_rgba = torch.empty((n_local, 4), dtype=torch.uint8).pin_memory() if want_rgba else None
        h_hits_ptr, h_rgba_ptr = h_hits.data_ptr(), (h_rgba.data_ptr() if want_rgba else None)
        e2e_bytes = n_local * 32 + (n_local * 4 if want_rgba else 0)
    for _ in range(3):
        scene.render_into(fh, h_hits_ptr, h_rgba_ptr)
    barrier()
    t0 = time.time()
    dbg = []
    for _ in range(args.steps):
        # camera matrices (the per-frame input, CameraUniformData) travel host->device as kernel parameters
        ta = time.time()
        scene.render_into(fh, h_hits_ptr, h_rgba_ptr)   # blocks until this rank's records have landed in host memory
        tb = time.time()
        if world > 1:
            dist.barrier()                               # the frame is complete when every rank's shard has landed
        dbg.append((1e3 * (tb - ta), 1e3 * (time.time() - tb), scene.timing()["trace_ms"]) if os.environ.get("VKHRT_BENCH_DEBUG") else None)
    if os.environ.get("VKHRT_BENCH_DEBUG"):
        print(f"rank {rank} e2e frames (render ms, barrier ms, kernel ms):", [tuple(round(x, 2) for x in d) for d in dbg], file=sys.stderr)
    torch.cuda.synchronize()
    e2e_wall = time.time() - t0
    e2e_t = torch.tensor([e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = rays_per_step * args.steps / float(e2e_t.item()) / 1e6
    frame_timing = scene.timing()
    # the same arm with TWO FRAMES IN FLIGHT (vkhrt_render_submit / _wait; the reference's Renderer keeps MAX_FRAMES_IN_FLIGHT frames behind
    # fences, renderer.cpp:85-119): frame k's records cross PCIe on the copy engine while frame k+1 traverses.  Every step still takes its
    # camera in and delivers its records to host memory inside the timed region.  Reported next to the blocking number, not instead of it.
    e2e_pipelined = None
    if world == 1 and shared_host is None:
        bufs = [(h_hits, h_rgba), (torch.empty_like(h_hits).pin_memory(), torch.empty_like(h_rgba).pin_memory() if want_rgba else None)]
        def submit(k):
            hb, ib = bufs[k % 2]
            scene.submit(fh, hb.data_ptr(), ib.data_ptr() if want_rgba else None)
        for k in range(3):
            submit(k)
        scene.wait(); scene.wait()
        torch.cuda.synchronize()
        t0 = time.time()
        for k in range(args.steps):
            if k >= 2:
                scene.wait()
            submit(k)
        for _ in range(min(2, args.steps)):
            scene.wait()
        e2e_pipelined = rays_per_step * args.steps / (time.time() - t0) / 1e6
        pipelined_ok = bool(torch.equal(bufs[0][0], bufs[1][0]))          # same camera: both buffer sets hold the same frame
    clocks = sampler.stop() if sampler else None        # sampled under load only: the device-timed and end-to-end arms
    e2e_frame_ok = None
    if shared_host is not None:
        barrier()
        if rank == 0:
            e2e_frame_ok = shared_host.hits().tobytes() == frame_hits.tobytes()    # the host frame == the device-assembled frame
        barrier()
        shared_host.close()

    # ---------------- algorithmic bytes per ray (GPU debug counters on the same frame) ----------------
    d_hits = torch.empty((n_local, 32), dtype=torch.uint8, device=dev)
    fs = V.make_frame(vi, pi, W, H, spp=spp, output_memory=V.MEM_DEVICE, stream=stream.cuda_stream, **sharded.layout.frame_kwargs(rank))
    stats = scene.render_stats_into(fs, d_hits.data_ptr(), None)
    torch.cuda.synchronize()
    agg = torch.tensor([stats["rays"], stats["nodes_visited"], stats["prims_tested"], stats["hits"], stats["phantom_iterations"]],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg)
    rays_c, nodes_c, prims_c, hits_c, iters_c = [float(x) for x in agg.tolist()]
    sched = {k: {"steps": int(s_), "lanes_per_step": (l_ / s_ if s_ else 0.0)}
             for k, s_, l_ in zip(("node", "leaf", "march", "refill"), stats["sched_steps"], stats["sched_lanes"])}

    # ---------------- the assembled frame against rank 0's own single-GPU render (N > 1, outside the timing) ----------------
    parity_assembled = None
    if world > 1:
        if rank == 0:
            whole = V.make_frame(vi, pi, W, H, spp=spp, output_memory=V.MEM_DEVICE, stream=stream.cuda_stream)
            d_whole = torch.empty((W * H, 32), dtype=torch.uint8, device=dev)
            d_whole_i = torch.empty((W * H, 4), dtype=torch.uint8, device=dev) if want_rgba else None
            scene.render_into(whole, d_whole.data_ptr(), d_whole_i.data_ptr() if want_rgba else None)
            torch.cuda.synchronize()
            n_tiles = sharded.layout.n_tiles
            parity_assembled = {"tiles_checked": n_tiles, "of_tiles": n_tiles,
                                "hits_identical": d_whole.cpu().numpy().tobytes() == frame_hits.tobytes(),
                                "against": "rank 0's own single-GPU render of the whole frame (every tile, every byte)"}
            if want_rgba:
                parity_assembled["rgba_identical"] = bool(np.array_equal(d_whole_i.cpu().numpy(), frame_rgba))
            if e2e_frame_ok is not None:
                parity_assembled["e2e_host_frame_identical"] = bool(e2e_frame_ok)
        dist.barrier()

    sharded.close()
    scene.close()
    del flush, d_hits
    torch.cuda.empty_cache()

    # ---------------- BASELINE configs[4] on the same N ranks: fixed 4K x 64 spp frame, strong scaling ----------------
    c5 = None
    if not args.no_strong_c5:
        c5 = strong_c5(V, torch, dist, world, rank, local_rank, dev, stream, max(2, args.c5_frames), 2)

    if rank == 0:
        peak, peak_src = peaks()
        out_bytes = (32.0 / spp if spp > 1 else 32.0) + (4.0 / spp if want_rgba else 0.0)
        n_int, n_prim = nodes_c / rays_c, prims_c / rays_c
        b_ray = 64.0 * n_int + PRIM_BYTES[tech] * n_prim + out_bytes
        # per GPU: every rank traces 1/N of the rays against its own copy of the scene, so the whole-job rate is divided by N
        achieved = value / world * 1e6 * b_ray / 1e9
        kernel_ms = device_timing["trace_ms"]
        cnt = profile_counters(name, tech) or {}
        traffic = cnt.get("dram_bytes_per_launch")
        line = {
            "metric": "Mrays/s primary-ray hair hits", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_desc(name, world, W, H, T), "rays_per_step": rays_per_step, "rays_per_gpu_per_step": rays_per_step // world,
                       "l2": "256 MB flush between timed frames; scene is %.0f MB (%d BVH leaves: a 64-byte node and a %d-byte primitive record each)" % (n_leaves * (64 + LEAF_RECORD_BYTES[tech]) / 1e6, n_leaves, LEAF_RECORD_BYTES[tech]),
                       "seed": hex(V.DEFAULT_SEED), "build_ms": build_timing["build_total_ms"],
                       "traversal_kernel": ("value: trace_pool_kernel (per-warp ray pool) for Phantom frames of >= 3x the pool's resident capacity, else "
                                            "trace_kernel (lane-bound); e2e: the same kernel delivering complete 128-byte lines of records to the page-locked host buffer"),
                       "frame_assembly": {"single": "one GPU", "peer": "traversal kernels store hit records straight into rank 0's frame buffer over NVLink (CUDA IPC peer mapping, two alternating buffer sets), 4-byte NCCL all_reduce as completion signal",
                                          "gather": "NCCL all_gather of compact shards + untile kernel"}[sharded.mode]},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 128 + 64,
                    "d2h_bytes_per_step": int(e2e_bytes) * (world if shared_host is not None else 1),
                    "pcie_gbs_per_rank": e2e_bytes * args.steps / float(e2e_t.item()) / 1e9,
                    "assembled_host_frame": shared_host is not None or world == 1,
                    "two_frames_in_flight": ({"value": e2e_pipelined, "unit": "Mrays/s", "frames_identical": pipelined_ok,
                                              "note": "vkhrt_render_submit / vkhrt_render_wait, 2 frames outstanding (the reference keeps frames in flight behind fences, "
                                                      "source/renderer.cpp:85-119): records leave on the copy engine while the next frame traverses"}
                                             if e2e_pipelined is not None else None),
                    "note": ("vkhrt_render with host buffers: camera in (kernel parameters), hit records out to page-locked host memory; wall clock."
                             + (" N > 1: every rank's kernel stores into ONE shared page-locked frame over its own PCIe link; the bound is the host side "
                                "(SM-issued 128-byte posted writes, ~44 GB/s per link measured with tools/micro/pcie_write.cu; links behind one PCIe switch / "
                                "NUMA node share it), not NVLink or NCCL" if shared_host is not None else ""))},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "per": "GPU (whole-job rate / n_gpus x B_ray)",
                         "bytes_per_ray": b_ray, "n_int_per_ray": n_int, "n_prim_per_ray": n_prim,
                         "phantom_iterations_per_ray": iters_c / rays_c, "hit_fraction": hits_c / rays_c,
                         "kernel_ms": kernel_ms, "kernel_ms_e2e_path": frame_timing["trace_ms"], "warp_scheduler_rank0": sched,
                         # what actually binds (ncu, profiles/): DRAM traffic vs peak, and the issue / lane budget
                         "dram_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "issue_active": cnt.get("issue_active"), "lanes_per_inst": cnt.get("lanes_per_inst"),
                         "regs": cnt.get("regs"), "warps_active": cnt.get("warps_active"), "l2_hit_rate": cnt.get("l2_hit_rate"),
                         "binding": cnt.get("binding"), "counters_source": cnt.get("source"),
                         "note": "frac = ALGORITHMIC bytes (B_ray = 64*N_int + P*N_prim + W from the GPU kernel's own debug counters) over the HBM peak: an "
                                 "upper bound the kernel would meet only if it were fetch-bound; the caches serve most of those bytes, so dram_frac "
                                 "(real DRAM traffic) and issue_active x lanes_per_inst/32 (the instruction budget used) say what binds (DESIGN.md §6)"},
        }
        if parity_assembled is not None:
            line["parity_assembled"] = parity_assembled
        orc = None
        if not args.no_parity or (not args.no_cpu_baseline and world == 1):
            orc, _ = oracle_scene(name)
        if not args.no_parity:
            from oracle.parity import stratified_pixels
            sub = stratified_pixels(W, H, N_PARITY)
            k = sub.astype(np.int64)
            line["parity"] = parity_block(orc, name, W, H, spp, want_rgba, sub, frame_hits[k], frame_rgba[k] if want_rgba else None)
        if not args.no_cpu_baseline and world == 1:      # the CPU leg runs at N=1 only
            from oracle.parity import stratified_pixels
            frame = oracle_frame(name, bw, bh)
            sub = stratified_pixels(bw, bh, bw * bh if name == "c1" else 262144)
            time_oracle(orc, frame, sub[:4096], 1)
            t1, _ = time_oracle(orc, frame, sub, 1)
            reps = int(max(1, min(256, args.cpu_seconds / max(t1[0], 1e-3))))
            tt, ost = time_oracle(orc, frame, sub, reps)
            cpu_mrays = len(sub) * reps / sum(tt) / 1e6
            line["cpu_baseline"] = {"value": cpu_mrays, "unit": "Mrays/s", "cores": host_threads(), "kind": "port",
                                    "sample": f"{reps} x {len(sub)} stratified pixels of the N=1 {bw}x{bh} frame ({sum(tt):.1f} s of CPU work)",
                                    "n_int_per_ray": ost["nodes_visited"] / ost["rays"], "n_prim_per_ray": ost["prims_tested"] / ost["rays"]}
        if c5 is not None:
            line["strong_c5"] = c5
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
