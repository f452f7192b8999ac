#!/usr/bin/env python
"""Device-time measurements of the rows around the traversal kernel (SURVEY.md §8 A1-A3, A10 and the (f) rows):
LBVH build, refit, strand LOD passes, shading with the constant / environment miss colour, ambient occlusion.
All times are CUDA-event times reported by the library (vkhrt_last_timing); bytes are ALGORITHMIC bytes per element
(stated per row below), so GB/s can be read against the measured HBM peak in MEASURED_PEAKS.json.

    python tools/bench_rows.py [--strands 100000] [--segments 32] [--reps 5] > profiles/r02_rows.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkhrt_b200 as V  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def row(name, ms, n, bytes_per, unit, note):
    gbs = n * bytes_per / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    return {"row": name, "ms": round(ms, 4), "elements": int(n), "unit": unit, "M_per_s": round(n / (ms * 1e3), 1) if ms > 0 else None,
            "algorithmic_bytes_per_element": bytes_per, "GB_per_s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak(), 4), "note": note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strands", type=int, default=100000)
    ap.add_argument("--segments", type=int, default=32)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    pos, idx = V.generate_groom(a.strands, a.segments, V.GROOM_CURLY)
    n = idx.shape[0]
    W, H = 1920, 1080
    vi, pi = V.camera_matrices(aspect=float(np.float32(W) / np.float32(H)))
    rows = []

    def best(fn):
        return min(fn() for _ in range(a.reps))

    # --- build / refit (Phantom, C2 groom) ---
    def build_once():
        with V.Scene(pos, idx) as sc:
            sc.build()
            return sc.timing()
    ts = [build_once() for _ in range(a.reps)]
    t = min(ts, key=lambda d: d["build_total_ms"])
    nl = 2 * n                       # BVH leaves: VKHRT_LEAF_SPLIT_PHANTOM = 2 pieces per curve
    rows.append(row("lbvh_build_total", t["build_total_ms"], nl, 64 + 28 + 8 * 24 + 8 + 24 + 208, "leaves",
                    "gen+centroid 64 B, morton 28 B, onesweep: 8 B histogram read + 8 passes x 24 B, karras 24 B write, materialise+refit 208 B"))
    rows.append(row("lbvh_geometry_centroids", t["geometry_ms"], nl, 64, "leaves", "curve generation from 4 vertices through the index pairs, piece box, centroid out (16 B)"))
    rows.append(row("lbvh_sort", t["sort_ms"], nl, 8 * 24 + 8, "keys", "onesweep: one histogram read (8 B) + 8 digit passes x (12 B read + 12 B write), decoupled look-back"))
    rows.append(row("lbvh_karras", t["hierarchy_ms"], nl, 24 + 16, "leaves", "neighbour keys read (L1/L2), child refs + parents + local bit written"))
    rows.append(row("lbvh_materialise_refit", t["refit_ms"], nl, 64 + 64 + 64 + 16, "leaves",
                    "4 vertices + indices in (~64 B), primA+primB out (64 B), node boxes written (64 B) and parents read (16 B)"))
    with V.Scene(pos, idx) as sc:
        sc.build()
        def refit():
            sc.refit(pos)
            return sc.timing()["refit_ms"]
        rows.append(row("refit_only", best(refit), nl, 64 + 64 + 64 + 16, "leaves", "vkhrt_scene_refit: materialise_refit_kernel + upper_refit_kernel (H2D of the positions not included)"))

        # --- shading rows at 1080p ---
        env = V.generate_environment(2048, 1024)
        sc.set_environment(env)
        def shade(**kw):
            def f():
                sc.render(V.make_frame(vi, pi, W, H, **kw), hits=False, rgba=True)
                return sc.timing()
            return f
        s0 = min((shade()() for _ in range(a.reps)), key=lambda d: d["shade_ms"])
        rows.append(row("shade_constant_miss", s0["shade_ms"], W * H, 36, "pixels", "32 B hit record in, 4 B RGBA8 out"))
        s1 = min((shade(miss_mode=V.MISS_ENVIRONMENT)() for _ in range(a.reps)), key=lambda d: d["shade_ms"])
        rows.append(row("shade_environment_miss", s1["shade_ms"], W * H, 36, "pixels",
                        "same + miss.rmiss for the ~14 % miss pixels (ray regenerated, 4 texels of a 2048x1024 RGBA32F map, asin/atan/exp/pow)"))
        s2 = min((shade(ao_samples=4)() for _ in range(a.reps)), key=lambda d: d["ao_ms"])
        hits = sc.render(V.make_frame(vi, pi, W, H), hits=True, rgba=False)[0]
        n_hit = int((hits["flags"] & 1).sum())
        rows.append(row("ambient_occlusion_4_rays", s2["ao_ms"], 4 * n_hit, 0, "AO rays", "terminate-on-first-hit rays through trace_kernel<ANYHIT>; Mrays/s in M_per_s"))

    # --- strand LOD passes on the same 3.2 M lines ---
    def lod(split, merge, cmerge):
        def f():
            with V.Scene(pos, idx) as sc:
                sc.apply_lod(split, merge, cmerge)
                sc.build()
                return sc.timing()["lod_ms"], sc.n_segments
        return f
    ms, ns = min(lod(1, 0, 0)() for _ in range(a.reps))
    rows.append(row("lod_split_lines_x1", ms, n, 36 + 32 + 64 + 64 + 72, "lines in",
                    "lines_from_mesh (36 B in, 32 B out) + SplitLines (32 in, 64 out) + lines_to_mesh (64 in, 2x(24+8+0) out); -> %d lines" % ns))
    ms, ns = min(lod(0, 1, 0)() for _ in range(a.reps))
    rows.append(row("lod_merge_lines_x1", ms, n, 36 + 32 + 16 + 2 + 32 + 16 + 16 + 36, "lines in",
                    "lines_from_mesh + count (16 B/line) + scan + scatter (32 in, ~16 out) + lines_to_mesh; 1 host read of the compacted size; -> %d lines" % ns))
    ms, ns = min(lod(0, 0, 1)() for _ in range(a.reps))
    rows.append(row("lod_merge_curves_fast_x1", ms, n, 36 + 32 + 68 + 48 + 48 + 24 + 36, "curves in",
                    "GenerateCurves materialised (48 B) + count + scan + scatter (48 in, ~24 out) + lines_from_curves + lines_to_mesh; -> %d curves" % ns))
    print(json.dumps({"device": "B200", "groom": f"curly {a.strands} x {a.segments}", "hbm_peak_GB_per_s": peak(), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
