#!/usr/bin/env python
"""bench.py — Mrays/s of the primary-ray hair-hit path (ray-gen + BVH traversal + intersection +
hit-record write) on B200, with the CPU oracle timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "c2"): synthetic curly groom 100k strands x 32 segments
(3.2 M Phantom curves), 1920x1080 primary rays per GPU, Phantom intersector, hit buffer only.
N > 1 (weak scaling): the BVH is replicated, the frame grows to N x 1080p pixels at fixed aspect and
camera, 64x64 tiles are dealt round-robin to ranks, shards are gathered with NCCL and untiled.
One step = one frame.  `value` is timed with CUDA events on the launching stream with outputs resident in
HBM; `e2e` goes through vkhrt_render with HOST buffers (camera in, hit records out to pinned memory).
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


def frame_size(base_w, base_h, n):
    if n == 1:
        return base_w, base_h
    w = int(math.ceil(base_w * math.sqrt(n) / 8.0) * 8)
    h = int(round(w * base_h / base_w))
    return w, h


def workload_desc(name, n, w, h):
    s, g, style, _, _, tech, spp, rgba = WORKLOADS[name]
    return (f"{name}: synthetic {style} groom {s} strands x {g} segments ({s * g} segs), {tech} intersector, "
            f"{w}x{h} primary rays x {spp} spp, {'hit buffer + RGBA8' if rgba else 'hit buffer only'}"
            + (f", 64x64 tiles round-robin over {n} GPUs" if n > 1 else ""))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
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
        sm, mx, reasons = [], [], set()
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def oracle_sample(name, n_rays, seed=0x5EED):
    """CPU leg: the oracle over a stratified pixel subset of the N=1 workload, all host threads."""
    import vkhrt_b200 as V
    from oracle import oracle as O
    s, g, style, w, h, tech, spp, rgba = WORKLOADS[name]
    tech_id = {"phantom": 0, "lss": 1, "dots": 2}[tech]
    pos, idx = V.generate_groom(s, g, V.GROOM_STRAIGHT if style == "straight" else V.GROOM_CURLY)
    vi, pi = V.camera_matrices(aspect=float(np.float32(w) / np.float32(h)))
    t0 = time.time()
    orc = O.OracleScene(pos, idx, technique=tech_id)
    build_s = time.time() - t0
    n_rays = min(n_rays, w * h)
    rng = np.random.default_rng(seed)
    # stratified: one random pixel from each of n_rays equal strata of the row-major frame
    edges = np.linspace(0, w * h, n_rays + 1).astype(np.int64)
    sub = (edges[:-1] + (rng.random(n_rays) * np.maximum(1, np.diff(edges))).astype(np.int64)).astype(np.uint64)
    frame = O.make_frame(vi, pi, w, h)
    return orc, frame, sub, build_s


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ask the affinity mask instead)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_oracle(orc, frame, sub, reps):
    times, stats = [], None
    for _ in range(reps):
        t0 = time.time()
        _, _, stats = orc.render(frame, hits=True, rgba=False, pixel_subset=sub, stats=True, n_threads=host_threads())
        times.append(time.time() - t0)
    return times, stats


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  vkhrt has none (Vulkan RT, Windows only)
    and cannot be compiled here, so this is the oracle port on all host threads (oracle/vkhrt_oracle.cpp)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    name = args.workload
    s, g, style, w, h, tech, spp, rgba = WORKLOADS[name]
    n_sample = 262144
    orc, frame, sub, build_s = oracle_sample(name, n_sample)
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
        "config": {"workload": workload_desc(name, 1, w, h), "sample": sample},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = mmzala/vkhrt has no CPU path and cannot be built here (Vulkan RT pipeline, Windows); "
                "this arm times the line-for-line C++ port of its shaders (oracle/), OpenMP over all host threads",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline duration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from vkhrt_b200.multi import ShardedRenderer
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    T = 64
    sharded = ShardedRenderer(scene, W, H, tile=T, spp=spp, want_rgba=want_rgba, device=dev, mode=args.gather)
    fd = sharded.make_frame(vi, pi, stream.cuda_stream)
    n_local = sharded.layout.shard_pixels
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_device():
        # trace this rank's tiles; N > 1: NCCL all_gather of the compact shards + untile (vkhrt_b200/multi.py)
        sharded.render(fd, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel-level arm: outputs stay in HBM ----------------
    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = V.launch_count()
    barrier()
    t_wall0 = time.time()
    for a, b in evs:
        flush.zero_()                 # evict the scene from L2 between timed frames (not inside the events)
        a.record(stream)
        step_device()
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

    # ---------------- end-to-end arm: the public call with HOST buffers ----------------
    fh = V.make_frame(vi, pi, W, H, spp=spp, tile_size=T, tile_first=rank, tile_stride=world, output_memory=V.MEM_HOST)
    h_hits = torch.empty((n_local, 32), dtype=torch.uint8).pin_memory()
    h_rgba = torch.empty((n_local, 4), dtype=torch.uint8).pin_memory() if want_rgba else None
    for _ in range(3):
        scene.render_into(fh, h_hits.data_ptr(), h_rgba.data_ptr() if want_rgba else None)
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        # camera matrices (the per-frame input, CameraUniformData) travel host->device as kernel parameters
        scene.render_into(fh, h_hits.data_ptr(), h_rgba.data_ptr() if want_rgba else None)   # blocks until the D2H copy landed
    barrier()
    e2e_t = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = rays_per_step * args.steps / float(e2e_t.item()) / 1e6
    clocks = sampler.stop() if sampler else None
    frame_timing = scene.timing()

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

    if rank == 0:
        peak, peak_src = peaks()
        out_bytes = (32.0 / spp if spp > 1 else 32.0) + (4.0 / spp if want_rgba else 0.0)
        n_int, n_prim = nodes_c / rays_c, prims_c / rays_c
        b_ray = 64.0 * n_int + PRIM_BYTES[tech] * n_prim + out_bytes
        achieved = value * 1e6 * b_ray / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(f"{name}_{tech}_bytes_per_launch")
        line = {
            "metric": "Mrays/s primary-ray hair hits", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_desc(name, world, W, H), "rays_per_step": rays_per_step, "rays_per_gpu_per_step": rays_per_step // world,
                       "l2": "256 MB flush between timed frames; scene is %.0f MB (%d BVH leaves: a 64-byte node and a %d-byte primitive record each)" % (n_leaves * (64 + LEAF_RECORD_BYTES[tech]) / 1e6, n_leaves, LEAF_RECORD_BYTES[tech]),
                       "seed": hex(V.DEFAULT_SEED), "build_ms": build_timing["build_total_ms"],
                       "traversal_kernel": ("value: trace_pool_kernel (per-warp ray pool) for Phantom frames of >= 3x the pool's resident capacity, else "
                                            "trace_kernel (lane-bound); e2e: the same kernel delivering complete 128-byte lines of records to the pinned host buffer"),
                       "frame_assembly": {"single": "one GPU", "peer": "traversal kernels store hit records straight into rank 0's frame buffer over NVLink (CUDA IPC peer mapping), 4-byte NCCL all_reduce as completion signal",
                                          "gather": "NCCL all_gather of compact shards + untile kernel"}[sharded.mode]},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 128 + 64,
                    "d2h_bytes_per_step": int(n_local * 32 + (n_local * 4 if want_rgba else 0)),
                    "note": "vkhrt_render with host buffers: camera in (kernel parameters), hit records out to pinned host memory; wall clock"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src,
                         "bytes_per_ray": b_ray, "n_int_per_ray": n_int, "n_prim_per_ray": n_prim,
                         "phantom_iterations_per_ray": iters_c / rays_c, "hit_fraction": hits_c / rays_c,
                         "kernel_ms": device_timing["trace_ms"], "kernel_ms_e2e_path": frame_timing["trace_ms"], "warp_scheduler_rank0": sched,
                         "note": "B_ray = 64*N_int + P*N_prim + W from the GPU kernel's own debug counters (L2-resident upper levels "
                                 "make this exceed DRAM traffic; see DESIGN.md §6)"},
        }
        if not args.no_cpu_baseline and world == 1:      # the CPU leg runs at N=1 only
            from oracle import oracle as O
            n0 = 262144
            orc, frame, sub, _ = oracle_sample(name, n0)
            time_oracle(orc, frame, sub[:4096], 1)
            t1, _ = time_oracle(orc, frame, sub, 1)
            reps = int(max(1, min(256, args.cpu_seconds / max(t1[0], 1e-3))))
            tt, ost = time_oracle(orc, frame, sub, reps)
            cpu_mrays = len(sub) * reps / sum(tt) / 1e6
            line["cpu_baseline"] = {"value": cpu_mrays, "unit": "Mrays/s", "cores": host_threads(), "kind": "port",
                                    "sample": f"{reps} x {len(sub)} stratified pixels of the N=1 {bw}x{bh} frame ({sum(tt):.1f} s of CPU work)",
                                    "n_int_per_ray": ost["nodes_visited"] / ost["rays"], "n_prim_per_ray": ost["prims_tested"] / ost["rays"]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
