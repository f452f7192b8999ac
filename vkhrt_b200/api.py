"""ctypes host binding of include/vkhrt_b200.h.

Mirrors the reference's host surface for the hot path:
  ModelLoader::LoadFromFile + ProcessHair{Curves,LSS,DOTS}  -> Scene(positions, indices, technique)
  BottomLevelAccelerationStructure ctor                     -> Scene.build()
  FlyCamera::ViewMatrix/ProjectionMatrix + UpdateCameraResource -> FlyCamera.matrices()
  Renderer::Render / vkCmdTraceRaysKHR(W,H,1)               -> Scene.render(frame)
No compute happens in Python and there is no fallback: a missing library raises ImportError,
a missing GPU raises VkhrtError(NO_DEVICE) from the first compute call.
"""
import ctypes as C
import os
import numpy as np

PHANTOM, LSS, DOTS = 0, 1, 2
SHADE, DEBUG_PRIMID, SHADE_MATERIAL = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
MISS_CONSTANT, MISS_ENVIRONMENT = 0, 1
GROOM_STRAIGHT, GROOM_CURLY = 0, 1
DEFAULT_SEED = 0x5EED0001
FLOATS_PER_PRIM = {PHANTOM: 12, LSS: 8, DOTS: 9}

HIT_DTYPE = np.dtype([("t", "<f4"), ("segment", "<u4"), ("u", "<f4"), ("nx", "<f4"), ("ny", "<f4"),
                      ("nz", "<f4"), ("primitive", "<u4"), ("flags", "<u4")])
NODE_DTYPE = np.dtype([("lo0", "<f4", 3), ("child0", "<u4"), ("hi0", "<f4", 3), ("child1", "<u4"),
                       ("lo1", "<f4", 3), ("prim0", "<u4"), ("hi1", "<f4", 3), ("prim1", "<u4")])

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_lib", "libvkhrt_b200.so")


class VkhrtError(RuntimeError):
    def __init__(self, status, where):
        self.status = int(status)
        L = lib()
        msg = L.vkhrt_error_string(self.status).decode()
        detail = L.vkhrt_last_error().decode()
        super().__init__(f"{where}: {msg} ({self.status}){': ' + detail if detail else ''}")


class SceneDesc(C.Structure):
    _fields_ = [("positions_xyz", C.c_void_p), ("n_vertices", C.c_uint32), ("line_indices", C.c_void_p),
                ("n_segments", C.c_uint32), ("radius_per_vertex", C.c_void_p), ("radius", C.c_float),
                ("technique", C.c_int32), ("device", C.c_int32)]


class FrameDesc(C.Structure):
    _fields_ = [("view_inverse", C.c_float * 16), ("proj_inverse", C.c_float * 16),
                ("width", C.c_uint32), ("height", C.c_uint32), ("t_min", C.c_float), ("t_max", C.c_float),
                ("spp", C.c_uint32), ("shade_mode", C.c_int32), ("miss_rgb", C.c_float * 3),
                ("tile_size", C.c_uint32), ("tile_first", C.c_uint32), ("tile_stride", C.c_uint32),
                ("row_major_output", C.c_uint32), ("output_memory", C.c_int32), ("stream", C.c_void_p),
                ("ao_samples", C.c_uint32), ("ao_distance", C.c_float), ("ao_bias", C.c_float), ("miss_mode", C.c_int32)]


class LineAsset(C.Structure):
    _fields_ = [("positions_xyz", C.c_void_p), ("n_vertices", C.c_uint32), ("line_indices", C.c_void_p),
                ("n_segments", C.c_uint32), ("radius_per_vertex", C.c_void_p), ("n_strands", C.c_uint32), ("base_color", C.c_float * 4)]


class Material(C.Structure):
    _fields_ = [("albedo_factor", C.c_float * 4), ("albedo_map_rgba32f", C.c_void_p), ("albedo_map_width", C.c_uint32), ("albedo_map_height", C.c_uint32)]


class BvhView(C.Structure):
    _fields_ = [("n_primitives", C.c_uint32), ("n_nodes", C.c_uint32), ("nodes", C.c_void_p),
                ("sorted_prim_ids", C.c_void_p), ("sorted_morton", C.c_void_p),
                ("scene_lo", C.c_float * 3), ("scene_hi", C.c_float * 3)]


class Timing(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("geometry_ms", "morton_ms", "sort_ms", "hierarchy_ms", "refit_ms",
                                         "build_total_ms", "raygen_ms", "trace_ms", "shade_ms",
                                         "render_total_ms", "h2d_ms", "d2h_ms", "ao_ms", "lod_ms")]

    def as_dict(self):
        return {k: float(getattr(self, k)) for k, _ in self._fields_}


class TraceStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("prims_tested", C.c_uint64),
                ("hits", C.c_uint64), ("phantom_iterations", C.c_uint64),
                ("sched_steps", C.c_uint64 * 4), ("sched_lanes", C.c_uint64 * 4)]

    def as_dict(self):
        d = {k: int(getattr(self, k)) for k, _ in self._fields_[:5]}
        d["sched_steps"] = [int(x) for x in self.sched_steps]
        d["sched_lanes"] = [int(x) for x in self.sched_lanes]
        return d


# every symbol include/vkhrt_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "vkhrt_abi_version", "vkhrt_device_count", "vkhrt_error_string", "vkhrt_last_error", "vkhrt_launch_count",
    "vkhrt_scene_create", "vkhrt_scene_build", "vkhrt_scene_refit", "vkhrt_scene_get_bvh",
    "vkhrt_scene_get_primitives", "vkhrt_scene_primitive_count", "vkhrt_scene_destroy",
    "vkhrt_render", "vkhrt_render_submit", "vkhrt_render_wait", "vkhrt_render_stats", "vkhrt_frame_local_pixels", "vkhrt_untile", "vkhrt_untile_host", "vkhrt_render_multi", "vkhrt_last_timing",
    "vkhrt_generate_rays", "vkhrt_trace_rays", "vkhrt_trace_rays_any_hit", "vkhrt_camera_matrices", "vkhrt_groom_generate",
    "vkhrt_host_alloc", "vkhrt_host_free", "vkhrt_shared_buffer_create", "vkhrt_shared_buffer_open", "vkhrt_shared_buffer_close", "vkhrt_shared_buffer_destroy",
    "vkhrt_scene_set_environment", "vkhrt_scene_set_material", "vkhrt_scene_set_meshes", "vkhrt_scene_set_mesh_material", "vkhrt_scene_mesh_count", "vkhrt_scene_mesh_of_segments", "vkhrt_image_save_exr", "vkhrt_scene_apply_lod", "vkhrt_scene_segment_count", "vkhrt_scene_get_lines",
    "vkhrt_asset_load_lines", "vkhrt_asset_save_lines", "vkhrt_asset_free", "vkhrt_image_load_hdr", "vkhrt_image_save_hdr",
    "vkhrt_image_free", "vkhrt_image_save_png", "vkhrt_environment_generate",
]

_lib = None


def library_path():
    return _LIB_PATH


def lib():
    """Load the CUDA library. There is deliberately no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C vkhrt_b200/csrc). vkhrt_b200 has no CPU fallback.")
    L = C.CDLL(_LIB_PATH)
    L.vkhrt_abi_version.restype = C.c_int
    L.vkhrt_device_count.restype = C.c_int
    L.vkhrt_error_string.restype = C.c_char_p
    L.vkhrt_error_string.argtypes = [C.c_int]
    L.vkhrt_last_error.restype = C.c_char_p
    L.vkhrt_launch_count.restype = C.c_uint64
    L.vkhrt_scene_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]
    L.vkhrt_scene_build.argtypes = [C.c_void_p]
    L.vkhrt_scene_refit.argtypes = [C.c_void_p, C.c_void_p]
    L.vkhrt_scene_get_bvh.argtypes = [C.c_void_p, C.POINTER(BvhView)]
    L.vkhrt_scene_get_primitives.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.vkhrt_scene_primitive_count.restype = C.c_uint32
    L.vkhrt_scene_primitive_count.argtypes = [C.c_void_p]
    L.vkhrt_scene_destroy.argtypes = [C.c_void_p]
    L.vkhrt_scene_destroy.restype = None
    L.vkhrt_render.argtypes = [C.c_void_p, C.POINTER(FrameDesc), C.c_void_p, C.c_void_p]
    L.vkhrt_render_submit.argtypes = [C.c_void_p, C.POINTER(FrameDesc), C.c_void_p, C.c_void_p]
    L.vkhrt_render_wait.argtypes = [C.c_void_p]
    L.vkhrt_render_stats.argtypes = [C.c_void_p, C.POINTER(FrameDesc), C.c_void_p, C.c_void_p, C.POINTER(TraceStats)]
    L.vkhrt_frame_local_pixels.restype = C.c_uint64
    L.vkhrt_frame_local_pixels.argtypes = [C.POINTER(FrameDesc)]
    L.vkhrt_untile.argtypes = [C.POINTER(FrameDesc), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.vkhrt_untile_host.argtypes = [C.POINTER(FrameDesc), C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]
    L.vkhrt_render_multi.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.POINTER(FrameDesc), C.c_void_p, C.c_void_p]
    L.vkhrt_last_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    L.vkhrt_generate_rays.argtypes = [C.POINTER(FrameDesc), C.c_uint32, C.c_void_p, C.c_int]
    L.vkhrt_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.vkhrt_trace_rays_any_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.vkhrt_camera_matrices.restype = None
    L.vkhrt_camera_matrices.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.vkhrt_groom_generate.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p]
    L.vkhrt_shared_buffer_create.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    L.vkhrt_shared_buffer_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    L.vkhrt_shared_buffer_close.argtypes = [C.c_int, C.c_void_p]
    L.vkhrt_shared_buffer_destroy.argtypes = [C.c_int, C.c_void_p]
    L.vkhrt_scene_set_environment.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.vkhrt_scene_set_material.argtypes = [C.c_void_p, C.POINTER(Material)]
    L.vkhrt_scene_set_meshes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.vkhrt_scene_set_mesh_material.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(Material)]
    L.vkhrt_scene_mesh_count.restype = C.c_uint32
    L.vkhrt_scene_mesh_count.argtypes = [C.c_void_p]
    L.vkhrt_scene_mesh_of_segments.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.vkhrt_image_save_exr.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.vkhrt_scene_apply_lod.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
    L.vkhrt_scene_segment_count.restype = C.c_uint32
    L.vkhrt_scene_segment_count.argtypes = [C.c_void_p]
    L.vkhrt_scene_get_lines.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.vkhrt_asset_load_lines.argtypes = [C.c_char_p, C.POINTER(LineAsset)]
    L.vkhrt_asset_save_lines.argtypes = [C.c_char_p, C.POINTER(LineAsset)]
    L.vkhrt_asset_free.argtypes = [C.POINTER(LineAsset)]
    L.vkhrt_asset_free.restype = None
    L.vkhrt_image_load_hdr.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.vkhrt_image_save_hdr.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.vkhrt_image_free.argtypes = [C.c_void_p]
    L.vkhrt_image_free.restype = None
    L.vkhrt_image_save_png.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.vkhrt_environment_generate.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
    L.vkhrt_environment_generate.restype = None
    L.vkhrt_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.vkhrt_host_free.argtypes = [C.c_void_p]
    if L.vkhrt_abi_version() != 6:
        raise ImportError("libvkhrt_b200.so ABI version mismatch")
    _lib = L
    return L


def _check(rc, where):
    if rc != 0:
        raise VkhrtError(rc, where)


def device_count():
    return int(lib().vkhrt_device_count())


def launch_count():
    return int(lib().vkhrt_launch_count())


def camera_matrices(position=(0.0, 150.0, 20.0), yaw=-90.0, pitch=0.0, fov=60.0, aspect=16.0 / 9.0, near=0.1, far=1000.0):
    """(view_inverse, proj_inverse) as float32[16] column-major; defaults = reference application.cpp:65-73."""
    pos = (C.c_float * 3)(*[float(x) for x in position])
    vi = (C.c_float * 16)()
    pi = (C.c_float * 16)()
    lib().vkhrt_camera_matrices(pos, yaw, pitch, fov, aspect, near, far, vi, pi)
    return np.array(list(vi), np.float32), np.array(list(pi), np.float32)


class FlyCamera:
    """Host mirror of the reference FlyCamera (source/fly_camera.cpp) without the input handling."""

    def __init__(self, position=(0.0, 150.0, 20.0), fov=60.0, aspect=16.0 / 9.0, near=0.1, far=1000.0, yaw=-90.0, pitch=0.0):
        self.position, self.fov, self.aspect, self.near, self.far, self.yaw, self.pitch = position, fov, aspect, near, far, yaw, pitch

    def matrices(self):
        return camera_matrices(self.position, self.yaw, self.pitch, self.fov, self.aspect, self.near, self.far)


def generate_groom(n_strands, segments, style=GROOM_CURLY, seed=DEFAULT_SEED):
    """Seeded synthetic groom -> (positions float32[n,3], line indices uint32[m,2]) in the Assimp line-mesh shape."""
    pos = np.empty((n_strands * (segments + 1), 3), np.float32)
    idx = np.empty((n_strands * segments, 2), np.uint32)
    _check(lib().vkhrt_groom_generate(n_strands, segments, style, seed, pos.ctypes.data, idx.ctypes.data), "vkhrt_groom_generate")
    return pos, idx


def make_frame(view_inv, proj_inv, width, height, spp=1, shade_mode=SHADE, miss_rgb=(0.0, 0.0, 0.0), tile_size=0,
               tile_first=0, tile_stride=0, t_min=0.0, t_max=0.0, output_memory=MEM_HOST, stream=None, row_major_output=0,
               ao_samples=0, ao_distance=0.0, ao_bias=0.0, miss_mode=MISS_CONSTANT):
    f = FrameDesc()
    f.ao_samples, f.ao_distance, f.ao_bias = int(ao_samples), float(ao_distance), float(ao_bias)
    f.miss_mode = int(miss_mode)
    f.view_inverse[:] = [float(x) for x in np.asarray(view_inv, np.float32).reshape(16)]
    f.proj_inverse[:] = [float(x) for x in np.asarray(proj_inv, np.float32).reshape(16)]
    f.width, f.height, f.spp, f.shade_mode = int(width), int(height), int(spp), int(shade_mode)
    f.t_min, f.t_max = t_min, t_max
    f.miss_rgb[:] = [float(x) for x in miss_rgb]
    f.tile_size, f.tile_first, f.tile_stride = tile_size, tile_first, tile_stride
    f.output_memory = output_memory
    f.stream = stream
    f.row_major_output = row_major_output
    return f


def load_material(path):
    """Material::albedoFactor of a line asset (glTF baseColorFactor of the first line primitive's material; 1,1,1,1 otherwise)"""
    a = LineAsset()
    _check(lib().vkhrt_asset_load_lines(os.fsencode(path), C.byref(a)), "vkhrt_asset_load_lines")
    try:
        return tuple(float(x) for x in a.base_color)
    finally:
        lib().vkhrt_asset_free(C.byref(a))


def load_lines(path):
    """ModelLoader::LoadFromFile for line assets (.gltf / .glb line primitives, .obj `l` records, .hair): -> (positions [n,3], indices [m,2], radius_per_vertex | None, n_strands)"""
    a = LineAsset()
    _check(lib().vkhrt_asset_load_lines(os.fsencode(path), C.byref(a)), "vkhrt_asset_load_lines")
    try:
        pos = np.ctypeslib.as_array(C.cast(a.positions_xyz, C.POINTER(C.c_float)), (a.n_vertices, 3)).copy() if a.n_vertices else np.zeros((0, 3), np.float32)
        idx = np.ctypeslib.as_array(C.cast(a.line_indices, C.POINTER(C.c_uint32)), (a.n_segments, 2)).copy() if a.n_segments else np.zeros((0, 2), np.uint32)
        rpv = None
        if a.radius_per_vertex and a.n_vertices:
            rpv = np.ctypeslib.as_array(C.cast(a.radius_per_vertex, C.POINTER(C.c_float)), (a.n_vertices,)).copy()
        return pos, idx, rpv, int(a.n_strands)
    finally:
        lib().vkhrt_asset_free(C.byref(a))


def save_lines(path, positions, indices, radius_per_vertex=None):
    pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    idx = np.ascontiguousarray(indices, np.uint32).reshape(-1, 2)
    rpv = None if radius_per_vertex is None else np.ascontiguousarray(radius_per_vertex, np.float32)
    a = LineAsset(pos.ctypes.data if pos.size else None, pos.shape[0], idx.ctypes.data if idx.size else None, idx.shape[0],
                  rpv.ctypes.data if rpv is not None else None, 0)
    _check(lib().vkhrt_asset_save_lines(os.fsencode(path), C.byref(a)), "vkhrt_asset_save_lines")


def load_hdr(path):
    """LoadFloatImageFromFile (stbi_loadf, 4 channels): Radiance .hdr -> float32 [h, w, 4]"""
    p = C.c_void_p(); w = C.c_uint32(); h = C.c_uint32()
    _check(lib().vkhrt_image_load_hdr(os.fsencode(path), C.byref(p), C.byref(w), C.byref(h)), "vkhrt_image_load_hdr")
    try:
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), (h.value, w.value, 4)).copy()
    finally:
        lib().vkhrt_image_free(p)


def save_hdr(path, rgba):
    e = np.ascontiguousarray(rgba, np.float32)
    _check(lib().vkhrt_image_save_hdr(os.fsencode(path), e.ctypes.data, e.shape[1], e.shape[0]), "vkhrt_image_save_hdr")


def save_exr(path, rgba):
    """float32 [h, w, 4] -> uncompressed scanline OpenEXR with FLOAT channels A, B, G, R"""
    e = np.ascontiguousarray(rgba, np.float32)
    _check(lib().vkhrt_image_save_exr(os.fsencode(path), e.ctypes.data, e.shape[1], e.shape[0]), "vkhrt_image_save_exr")


def save_png(path, rgba8, width, height):
    img = np.ascontiguousarray(rgba8, np.uint8).reshape(height, width, 4)
    _check(lib().vkhrt_image_save_png(os.fsencode(path), img.ctypes.data, width, height), "vkhrt_image_save_png")


def merge_meshes(meshes):
    """Concatenate line meshes [(positions [n, 3], indices [m, 2]) or (positions, indices, radius_per_vertex), ...] the way the reference's
    GenerateLines addresses them (firstVertex / firstIndex, geometry_processor.cpp:45-67): returns positions, indices (rebased),
    radius_per_vertex (None unless every mesh has one) and first_segment for Scene.set_meshes."""
    pos, idx, rad, first = [], [], [], []
    v0 = s0 = 0
    for m in meshes:
        p = np.ascontiguousarray(m[0], np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(m[1], np.uint32).reshape(-1, 2)
        pos.append(p); idx.append(i + np.uint32(v0)); first.append(s0)
        rad.append(None if len(m) < 3 or m[2] is None else np.ascontiguousarray(m[2], np.float32).reshape(-1))
        v0 += p.shape[0]; s0 += i.shape[0]
    if any(r is None for r in rad) and not all(r is None for r in rad):
        raise ValueError("merge_meshes: radius_per_vertex for all meshes or for none")
    radius = None if (not rad or rad[0] is None) else np.concatenate(rad)
    return (np.concatenate(pos) if pos else np.zeros((0, 3), np.float32), np.concatenate(idx) if idx else np.zeros((0, 2), np.uint32),
            radius, np.asarray(first, np.uint32))


def generate_environment(width=512, height=256):
    """procedural equirectangular sky, float32 [h, w, 4] (the reference's .hdr asset is not in its repository)"""
    out = np.empty((height, width, 4), np.float32)
    lib().vkhrt_environment_generate(width, height, out.ctypes.data)
    return out


def frame_local_pixels(frame):
    return int(lib().vkhrt_frame_local_pixels(C.byref(frame)))


def untile(frame, world, gathered_ptr, out_ptr, elem_bytes, stream=None):
    _check(lib().vkhrt_untile(C.byref(frame), world, gathered_ptr, out_ptr, elem_bytes, stream), "vkhrt_untile")


def untile_host(frame, world, gathered):
    """vkhrt_untile_host: compact shards concatenated rank-major -> row-major.  gathered: HIT_DTYPE[world * shard] or uint8[world * shard, 4]"""
    g = np.ascontiguousarray(gathered)
    n = frame.width * frame.height
    if g.dtype == HIT_DTYPE:
        out, elem = np.zeros(n, HIT_DTYPE), 32
    elif g.dtype == np.uint8 and g.ndim == 2 and g.shape[1] == 4:
        out, elem = np.zeros((n, 4), np.uint8), 4
    else:
        raise ValueError("untile_host: HIT_DTYPE records or [n, 4] uint8 pixels")
    _check(lib().vkhrt_untile_host(C.byref(frame), world, g.ctypes.data, out.ctypes.data, elem), "vkhrt_untile_host")
    return out


def render_multi(scenes, frame, hits=True, rgba=True):
    """vkhrt_render_multi: one frame on several GPUs from this process (scenes[r] = the same groom built on device r) ->
    (hits[H*W], rgba[H*W, 4]) in row-major order, identical to Scene.render of the whole frame"""
    frame.output_memory = MEM_HOST
    n = frame.width * frame.height
    h = np.zeros(n, HIT_DTYPE) if hits else None
    img = np.zeros((n, 4), np.uint8) if rgba else None
    arr = (C.c_void_p * len(scenes))(*[sc._h for sc in scenes])
    _check(lib().vkhrt_render_multi(arr, len(scenes), C.byref(frame), h.ctypes.data if hits else None, img.ctypes.data if rgba else None),
           "vkhrt_render_multi")
    return h, img


class Scene:
    """One groom on one GPU: primitives + LBVH resident in HBM."""

    def __init__(self, positions, indices, technique=PHANTOM, radius=0.02, radius_per_vertex=None, device=0):
        self._h = None
        pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, np.uint32).reshape(-1, 2)
        rpv = None if radius_per_vertex is None else np.ascontiguousarray(radius_per_vertex, np.float32)
        if rpv is not None and rpv.shape[0] != pos.shape[0]:
            raise ValueError("radius_per_vertex must have one entry per vertex")
        d = SceneDesc(pos.ctypes.data if pos.size else None, pos.shape[0], idx.ctypes.data if idx.size else None, idx.shape[0],
                      rpv.ctypes.data if rpv is not None else None, float(radius), int(technique), int(device))
        h = C.c_void_p()
        _check(lib().vkhrt_scene_create(C.byref(d), C.byref(h)), "vkhrt_scene_create")
        self._h = h
        self.technique = int(technique)
        self.n_segments = idx.shape[0]

    def close(self):
        if getattr(self, "_h", None):
            lib().vkhrt_scene_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def n_primitives(self):
        return int(lib().vkhrt_scene_primitive_count(self._h))

    @property
    def n_leaves(self):
        """BVH leaves = primitive groups x VKHRT_LEAF_SPLIT_* pieces (include/vkhrt_b200.h); read from the built scene"""
        v = BvhView()
        _check(lib().vkhrt_scene_get_bvh(self._h, C.byref(v)), "vkhrt_scene_get_bvh")
        return int(v.n_primitives)

    def apply_lod(self, line_split_passes=0, line_merge_passes=0, curve_merge_passes=0):
        """SplitLines / MergeLines / MergeCurvesFast on the device, before build()"""
        _check(lib().vkhrt_scene_apply_lod(self._h, line_split_passes, line_merge_passes, curve_merge_passes), "vkhrt_scene_apply_lod")
        self.n_segments = int(lib().vkhrt_scene_segment_count(self._h))
        return self

    def lines(self):
        out = np.empty((int(lib().vkhrt_scene_segment_count(self._h)), 6), np.float32)
        _check(lib().vkhrt_scene_get_lines(self._h, out.ctypes.data, out.size), "vkhrt_scene_get_lines")
        return out

    def set_environment(self, rgba):
        """RGBA32F equirectangular map [h, w, 4] for miss_mode=MISS_ENVIRONMENT (None removes it)"""
        if rgba is None:
            _check(lib().vkhrt_scene_set_environment(self._h, None, 0, 0), "vkhrt_scene_set_environment")
            return self
        e = np.ascontiguousarray(rgba, np.float32)
        if e.ndim != 3 or e.shape[2] != 4:
            raise ValueError("environment map must be [h, w, 4] float32")
        _check(lib().vkhrt_scene_set_environment(self._h, e.ctypes.data, e.shape[1], e.shape[0]), "vkhrt_scene_set_environment")
        return self

    def set_material(self, albedo_factor=(1.0, 1.0, 1.0, 1.0), albedo_map=None):
        """Material::albedoFactor and an optional RGBA32F albedo map [h, w, 4]; frames use it with shade_mode = SHADE_MATERIAL"""
        m = Material()
        m.albedo_factor[:] = [float(x) for x in albedo_factor]
        tex = None
        if albedo_map is not None:
            tex = np.ascontiguousarray(albedo_map, np.float32)
            if tex.ndim != 3 or tex.shape[2] != 4:
                raise ValueError("albedo map must be [h, w, 4] float32")
            m.albedo_map_rgba32f, m.albedo_map_width, m.albedo_map_height = tex.ctypes.data, tex.shape[1], tex.shape[0]
        _check(lib().vkhrt_scene_set_material(self._h, C.byref(m)), "vkhrt_scene_set_material")
        return self

    def set_meshes(self, first_segment):
        """Multi-mesh scene: mesh m owns segments [first_segment[m], first_segment[m + 1]) of the concatenated line list (see merge_meshes)"""
        fs = np.ascontiguousarray(first_segment, np.uint32).reshape(-1)
        _check(lib().vkhrt_scene_set_meshes(self._h, fs.ctypes.data if fs.size else None, fs.size), "vkhrt_scene_set_meshes")
        return self

    def set_mesh_material(self, mesh, albedo_factor=(1.0, 1.0, 1.0, 1.0), albedo_map=None):
        m = Material()
        m.albedo_factor[:] = [float(x) for x in albedo_factor]
        tex = None
        if albedo_map is not None:
            tex = np.ascontiguousarray(albedo_map, np.float32)
            if tex.ndim != 3 or tex.shape[2] != 4:
                raise ValueError("albedo map must be [h, w, 4] float32")
            m.albedo_map_rgba32f, m.albedo_map_width, m.albedo_map_height = tex.ctypes.data, tex.shape[1], tex.shape[0]
        _check(lib().vkhrt_scene_set_mesh_material(self._h, int(mesh), C.byref(m)), "vkhrt_scene_set_mesh_material")
        return self

    @property
    def n_meshes(self):
        return int(lib().vkhrt_scene_mesh_count(self._h))

    def mesh_of_segments(self, segments):
        """the mesh (= gl_InstanceCustomIndexEXT of the reference's TLAS) owning each segment index of a hit record; 0xFFFFFFFF for a miss"""
        seg = np.ascontiguousarray(segments, np.uint32)
        out = np.empty(seg.shape, np.uint32)
        _check(lib().vkhrt_scene_mesh_of_segments(self._h, seg.ctypes.data, out.ctypes.data, seg.size), "vkhrt_scene_mesh_of_segments")
        return out

    def build(self):
        _check(lib().vkhrt_scene_build(self._h), "vkhrt_scene_build")
        return self

    def refit(self, positions):
        pos = np.ascontiguousarray(positions, np.float32)
        _check(lib().vkhrt_scene_refit(self._h, pos.ctypes.data), "vkhrt_scene_refit")

    def bvh(self):
        v = BvhView()
        _check(lib().vkhrt_scene_get_bvh(self._h, C.byref(v)), "vkhrt_scene_get_bvh")
        nodes = np.zeros(v.n_nodes, NODE_DTYPE)
        ids = np.zeros(v.n_primitives, np.uint32)
        morton = np.zeros(v.n_primitives, np.uint64)
        v.nodes, v.sorted_prim_ids, v.sorted_morton = nodes.ctypes.data, ids.ctypes.data, morton.ctypes.data
        _check(lib().vkhrt_scene_get_bvh(self._h, C.byref(v)), "vkhrt_scene_get_bvh")
        lohi = np.array(list(v.scene_lo) + list(v.scene_hi), np.float32)
        return nodes, ids, morton, lohi

    def primitives(self):
        out = np.empty((self.n_primitives, FLOATS_PER_PRIM[self.technique]), np.float32)
        _check(lib().vkhrt_scene_get_primitives(self._h, out.ctypes.data, out.size), "vkhrt_scene_get_primitives")
        return out

    def render(self, frame, hits=True, rgba=True, stats=False):
        """Host-buffer render (the e2e path): returns (hits, rgba[, stats]) as numpy arrays."""
        n = frame_local_pixels(frame)
        frame.output_memory = MEM_HOST
        h = np.zeros(n, HIT_DTYPE) if hits else None
        img = np.zeros((n, 4), np.uint8) if rgba else None
        hp = h.ctypes.data if hits else None
        ip = img.ctypes.data if rgba else None
        if stats:
            st = TraceStats()
            _check(lib().vkhrt_render_stats(self._h, C.byref(frame), hp, ip, C.byref(st)), "vkhrt_render_stats")
            return h, img, st.as_dict()
        _check(lib().vkhrt_render(self._h, C.byref(frame), hp, ip), "vkhrt_render")
        return h, img, None

    def render_into(self, frame, hits_ptr=None, rgba_ptr=None):
        """Raw-pointer render (host or device pointers per frame.output_memory)."""
        _check(lib().vkhrt_render(self._h, C.byref(frame), hits_ptr, rgba_ptr), "vkhrt_render")

    def submit(self, frame, hits_ptr=None, rgba_ptr=None):
        """vkhrt_render_submit: enqueue a frame into HOST buffers (page-locked for an asynchronous copy) and return; up to 2 in flight"""
        _check(lib().vkhrt_render_submit(self._h, C.byref(frame), hits_ptr, rgba_ptr), "vkhrt_render_submit")

    def wait(self):
        """vkhrt_render_wait: the oldest outstanding frame's outputs are complete"""
        _check(lib().vkhrt_render_wait(self._h), "vkhrt_render_wait")

    def render_stats_into(self, frame, hits_ptr=None, rgba_ptr=None):
        st = TraceStats()
        _check(lib().vkhrt_render_stats(self._h, C.byref(frame), hits_ptr, rgba_ptr, C.byref(st)), "vkhrt_render_stats")
        return st.as_dict()

    def trace_rays(self, rays_ptr, n_rays, hits_ptr, stream=None, any_hit=False):
        """Wavefront trace of a device ray buffer; any_hit = terminate on the first accepted hit (occlusion rays)."""
        if any_hit:
            _check(lib().vkhrt_trace_rays_any_hit(self._h, rays_ptr, n_rays, hits_ptr, stream), "vkhrt_trace_rays_any_hit")
        else:
            _check(lib().vkhrt_trace_rays(self._h, rays_ptr, n_rays, hits_ptr, stream), "vkhrt_trace_rays")

    def timing(self):
        t = Timing()
        _check(lib().vkhrt_last_timing(self._h, C.byref(t)), "vkhrt_last_timing")
        return t.as_dict()


class HostBuffer:
    """Page-locked host memory from vkhrt_host_alloc, viewed as a numpy array (`.array`): hit-record buffers from here take the
    zero-copy path of vkhrt_render / vkhrt_render_multi."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _check(lib().vkhrt_host_alloc(max(n, 1), C.byref(p)), "vkhrt_host_alloc")
        self.ptr = p.value
        self.array = np.frombuffer((C.c_uint8 * max(n, 1)).from_address(self.ptr), dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            lib().vkhrt_host_free(self.ptr)
            self.ptr = None

    __del__ = close


class SharedBuffer:
    """A device buffer that the per-GPU processes of one box can all address (CUDA IPC; peer stores go over NVLink).
    The creator owns it: `SharedBuffer.create(bytes, device)` -> `.handle` (64 bytes) -> `SharedBuffer.open(handle, device)`."""

    def __init__(self, ptr, handle, device, owner):
        self.ptr, self.handle, self.device, self.owner = ptr, handle, device, owner

    @classmethod
    def create(cls, nbytes, device=0):
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        _check(lib().vkhrt_shared_buffer_create(device, nbytes, C.byref(ptr), handle), "vkhrt_shared_buffer_create")
        return cls(ptr.value, handle.raw, device, True)

    @classmethod
    def open(cls, handle, device=0):
        ptr = C.c_void_p()
        _check(lib().vkhrt_shared_buffer_open(device, bytes(handle), C.byref(ptr)), "vkhrt_shared_buffer_open")
        return cls(ptr.value, bytes(handle), device, False)

    def close(self):
        if self.ptr:
            if self.owner:
                lib().vkhrt_shared_buffer_destroy(self.device, self.ptr)
            else:
                lib().vkhrt_shared_buffer_close(self.device, self.ptr)
            self.ptr = None


def generate_rays(frame, sample, rays_ptr, device=0):
    _check(lib().vkhrt_generate_rays(C.byref(frame), sample, rays_ptr, device), "vkhrt_generate_rays")
