#!/bin/bash
# compute-sanitizer memcheck + racecheck over a small render of every technique (build + trace + shade kernels)
mkdir -p gpurun_out
cat > san_tmp.py <<'PY'
import numpy as np, vkhrt_b200 as V
pos, idx = V.generate_groom(1500, 12, V.GROOM_CURLY)
vi, pi = V.camera_matrices(aspect=float(np.float32(160) / np.float32(96)))
for tech in (V.PHANTOM, V.LSS, V.DOTS):
    with V.Scene(pos, idx, technique=tech) as sc:
        sc.build()
        h, img, st = sc.render(V.make_frame(vi, pi, 160, 96, spp=2), stats=False)
        sc.refit(pos + np.float32(0.01))
        h2, _, _ = sc.render(V.make_frame(vi, pi, 160, 96, tile_size=32, tile_first=1, tile_stride=2))
        print(tech, int((h["flags"] & 1).sum()), int((h2["flags"] & 1).sum()))
# strand LOD passes, environment miss shader, ambient occlusion
env = V.generate_environment(64, 32)
with V.Scene(pos, idx) as sc:
    sc.apply_lod(1, 1, 1).set_environment(env).build()
    h, img, _ = sc.render(V.make_frame(vi, pi, 160, 96, miss_mode=V.MISS_ENVIRONMENT, ao_samples=2))
    print("lod+env+ao", sc.n_segments, int((h["flags"] & 1).sum()), int(img[:, :3].sum()))
# hit records into a pinned host buffer (zero-copy stores; with the pool kernel: line-wise delivery), odd record count
import torch
with V.Scene(pos, idx) as sc:
    sc.build()
    hh = torch.zeros((161 * 97, 32), dtype=torch.uint8).pin_memory()
    sc.render_into(V.make_frame(vi, pi, 161, 97, output_memory=V.MEM_HOST), hh.data_ptr(), None)
    print("pinned", int(hh.numpy().view(V.HIT_DTYPE)["flags"].sum()))
# per-vertex radii (TAPER kernels), material shading, several samples per launch (spp 4 on a small frame)
taper = np.tile(np.linspace(0.02, 0.005, 13, dtype=np.float32), 1500)
for tech in (V.PHANTOM, V.DOTS):
    with V.Scene(pos, idx, technique=tech, radius_per_vertex=taper) as sc:
        sc.set_material((0.5, 0.6, 0.7, 1.0)).build()
        h, img, _ = sc.render(V.make_frame(vi, pi, 160, 96, spp=4, shade_mode=V.SHADE_MATERIAL))
        print("taper", tech, int((h["flags"] & 1).sum()), int(img[:, :3].sum()))
# three meshes in one scene: per-mesh materials through the mesh table
p2, i2 = V.generate_groom(400, 6, V.GROOM_STRAIGHT, seed=3)
mp, mi, _, first = V.merge_meshes([(pos, idx), (p2 + np.float32([4, 0, 0]), i2), (p2 - np.float32([4, 0, 0]), i2)])
with V.Scene(mp, mi) as sc:
    sc.set_meshes(first).build()
    sc.set_mesh_material(1, (0.2, 0.9, 0.4, 1.0)).set_mesh_material(2, (0.9, 0.2, 0.4, 1.0))
    h, img, _ = sc.render(V.make_frame(vi, pi, 160, 96, spp=2, shade_mode=V.SHADE_MATERIAL))
    print("meshes", np.bincount(sc.mesh_of_segments(h["segment"][(h["flags"] & 1) != 0]), minlength=3), int(img[:, :3].sum()))
# one process, two scene handles on one device: vkhrt_render_multi (peer path: shared frame buffers, persistent workers)
scs = [V.Scene(pos, idx, technique=V.LSS).build() for _ in range(2)]
h, img = V.render_multi(scs, V.make_frame(vi, pi, 160, 96, spp=2))
print("multi", int((h["flags"] & 1).sum()))
for s_ in scs:
    s_.close()
PY
# VKHRT_POOL_MIN_RATIO=0: the Phantom frames above go through the per-warp ray-pool kernel as well as the lane-bound one
for tool in memcheck racecheck; do
  for pool in 3 0; do
    VKHRT_POOL_MIN_RATIO=$pool timeout 900 compute-sanitizer --tool $tool --print-limit ${PRINT_LIMIT:-5} python san_tmp.py > gpurun_out/sanitizer_${tool}_pool$pool.log 2>&1
    echo "$tool (VKHRT_POOL_MIN_RATIO=$pool) rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_${tool}_pool$pool.log | tail -3
  done
done
rm -f san_tmp.py
