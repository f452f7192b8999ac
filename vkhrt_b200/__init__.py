"""vkhrt_b200 — B200-native hair ray-tracing hot path (CUDA sm_100a behind a C ABI).

The product is `_lib/libvkhrt_b200.so` (sources in `csrc/`, ABI in `include/vkhrt_b200.h`).
This package is the thin Python host binding used by tests and bench.py; it contains no
compute and no CPU fallback: importing `vkhrt_b200.api` without the built library raises.
"""
from .api import (  # noqa: F401
    PHANTOM, LSS, DOTS, SHADE, DEBUG_PRIMID, SHADE_MATERIAL, MEM_HOST, MEM_DEVICE, GROOM_STRAIGHT, GROOM_CURLY,
    HIT_DTYPE, NODE_DTYPE, FLOATS_PER_PRIM, DEFAULT_SEED,
    VkhrtError, FrameDesc, Scene, FlyCamera, camera_matrices, generate_groom, make_frame,
    frame_local_pixels, device_count, launch_count, library_path, untile, untile_host, render_multi, lib, HostBuffer, SharedBuffer,
    MISS_CONSTANT, MISS_ENVIRONMENT, load_lines, load_material, save_lines, load_hdr, save_hdr, save_png, save_exr, generate_environment, merge_meshes,
)
