"""ctypes binding of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module (as the checker / reported CPU baseline, never as the product path).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

HIT_DTYPE = np.dtype([("t", "<f4"), ("segment", "<u4"), ("u", "<f4"), ("nx", "<f4"), ("ny", "<f4"),
                      ("nz", "<f4"), ("primitive", "<u4"), ("flags", "<u4")])
NODE_DTYPE = np.dtype([("lo0", "<f4", 3), ("child0", "<u4"), ("hi0", "<f4", 3), ("child1", "<u4"),
                       ("lo1", "<f4", 3), ("prim0", "<u4"), ("hi1", "<f4", 3), ("prim1", "<u4")])
FLOATS_PER_PRIM = {0: 12, 1: 8, 2: 9}


class FrameDesc(C.Structure):
    _fields_ = [("view_inverse", C.c_float * 16), ("proj_inverse", C.c_float * 16),
                ("width", C.c_uint32), ("height", C.c_uint32), ("t_min", C.c_float), ("t_max", C.c_float),
                ("spp", C.c_uint32), ("shade_mode", C.c_int32), ("miss_rgb", C.c_float * 3),
                ("tile_size", C.c_uint32), ("tile_first", C.c_uint32), ("tile_stride", C.c_uint32),
                ("row_major_output", C.c_uint32), ("output_memory", C.c_int32), ("stream", C.c_void_p),
                ("ao_samples", C.c_uint32), ("ao_distance", C.c_float), ("ao_bias", C.c_float), ("miss_mode", C.c_int32)]


class TraceStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("prims_tested", C.c_uint64),
                ("hits", C.c_uint64), ("phantom_iterations", C.c_uint64),
                ("sched_steps", C.c_uint64 * 4), ("sched_lanes", C.c_uint64 * 4)]

    def as_dict(self):
        d = {k: int(getattr(self, k)) for k, _ in self._fields_[:5]}
        d["sched_steps"] = [int(x) for x in self.sched_steps]
        d["sched_lanes"] = [int(x) for x in self.sched_lanes]
        return d


def build(force=False):
    src = os.path.join(_HERE, "vkhrt_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "vkhrt_b200.h")
    incs = [os.path.join(_HERE, f) for f in ("vkhrt_oracle_studies_traversal.inc", "vkhrt_oracle_studies_api.inc")]
    if (force or not os.path.exists(_LIB_PATH)
            or any(os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in [src, hdr] + incs)):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "all"] if force else ["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_float, C.c_int]
        L.orc_scene_create_lod.restype = C.c_void_p
        L.orc_scene_create_lod.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.orc_scene_set_environment.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_scene_set_material.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_scene_set_meshes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_scene_set_mesh_material.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_scene_mesh_of_segment.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_scene_mesh_of_segment.restype = C.c_uint32
        L.orc_scene_line_count.restype = C.c_uint32
        L.orc_scene_line_count.argtypes = [C.c_void_p]
        L.orc_scene_get_lines.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_environment_miss.argtypes = [C.c_void_p, fp, fp]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_primitive_count.restype = C.c_uint32
        L.orc_scene_primitive_count.argtypes = [C.c_void_p]
        L.orc_scene_leaf_count.restype = C.c_uint32
        L.orc_scene_leaf_count.argtypes = [C.c_void_p]
        L.orc_scene_node_count.restype = C.c_uint32
        L.orc_scene_node_count.argtypes = [C.c_void_p]
        L.orc_scene_get_primitives.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_scene_get_aabbs.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_curve_aabb.argtypes = [fp, C.c_float, fp]
        L.orc_scene_leaf_split.restype = C.c_uint32
        L.orc_scene_leaf_split.argtypes = [C.c_void_p]
        L.orc_scene_get_bvh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_render.restype = C.c_int
        L.orc_render.argtypes = [C.c_void_p, C.POINTER(FrameDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                 C.c_int, C.POINTER(TraceStats), C.c_int]
        L.orc_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_ao_direction.argtypes = [fp, C.c_uint32, C.c_uint32, C.c_uint32, fp]
        L.orc_prhi.restype = C.c_int
        L.orc_prhi.argtypes = [fp, fp, fp, C.c_float, fp, fp, fp]
        L.orc_prhi_taper.restype = C.c_int
        L.orc_prhi_taper.argtypes = [fp, fp, fp, C.c_float, C.c_float, fp, fp, fp]
        L.orc_ray_cylinder.restype = C.c_int
        L.orc_ray_cylinder.argtypes = [fp, fp, fp, fp, C.c_float]
        L.orc_lss.restype = C.c_int
        L.orc_lss.argtypes = [fp, fp, fp, fp, fp, fp]
        L.orc_tri.restype = C.c_int
        L.orc_tri.argtypes = [fp, fp, fp, C.c_uint, fp, fp, fp]
        L.orc_raygen.argtypes = [fp, fp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, fp, fp]
        L.orc_curve_point.argtypes = [fp, C.c_float, fp]
        L.orc_curve_axis.argtypes = [fp, C.c_float, fp]
        L.orc_shade.argtypes = [fp, C.c_uint32, C.c_int, fp]
        L.orc_max_threads.restype = C.c_int
        L.orc_groom_generate.restype = C.c_int
        L.orc_groom_generate.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_camera_matrices.restype = None
        L.orc_camera_matrices.argtypes = [fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, fp, fp]
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


GROOM_STRAIGHT, GROOM_CURLY = 0, 1
DEFAULT_SEED = 0x5EED0001


def generate_groom(n_strands, segments, style=GROOM_CURLY, seed=DEFAULT_SEED):
    """The checker's own seeded groom; bit-identical to the product's vkhrt_groom_generate
    (tests/test_oracle_pins.py::test_oracle_inputs_equal_the_products)."""
    pos = np.empty((n_strands * (segments + 1), 3), np.float32)
    idx = np.empty((n_strands * segments, 2), np.uint32)
    rc = lib().orc_groom_generate(n_strands, segments, style, seed, pos.ctypes.data, idx.ctypes.data)
    if rc != 0:
        raise ValueError("orc_groom_generate: bad arguments")
    return pos, idx


def camera_matrices(position=(0.0, 150.0, 20.0), yaw=-90.0, pitch=0.0, fov=60.0, aspect=16.0 / 9.0, near=0.1, far=1000.0):
    """(view_inverse, proj_inverse) float32[16] column-major; defaults = reference application.cpp:65-73."""
    pos = (C.c_float * 3)(*[float(x) for x in position])
    vi = (C.c_float * 16)()
    pi = (C.c_float * 16)()
    lib().orc_camera_matrices(pos, yaw, pitch, fov, aspect, near, far, vi, pi)
    return np.array(list(vi), np.float32), np.array(list(pi), np.float32)


def max_threads():
    return int(lib().orc_max_threads())


def make_frame(view_inv, proj_inv, width, height, spp=1, shade_mode=0, miss_rgb=(0.0, 0.0, 0.0),
               tile_size=0, tile_first=0, tile_stride=0, t_min=0.0, t_max=0.0, row_major_output=0,
               ao_samples=0, ao_distance=0.0, ao_bias=0.0, miss_mode=0):
    f = FrameDesc()
    f.ao_samples, f.ao_distance, f.ao_bias = int(ao_samples), float(ao_distance), float(ao_bias)
    f.miss_mode = int(miss_mode)
    f.view_inverse[:] = [float(x) for x in np.asarray(view_inv, np.float32).reshape(16)]
    f.proj_inverse[:] = [float(x) for x in np.asarray(proj_inv, np.float32).reshape(16)]
    f.width, f.height, f.spp, f.shade_mode = width, height, spp, shade_mode
    f.t_min, f.t_max = t_min, t_max
    f.miss_rgb[:] = list(miss_rgb)
    f.tile_size, f.tile_first, f.tile_stride = tile_size, tile_first, tile_stride
    f.row_major_output = row_major_output
    return f


class OracleScene:
    def __init__(self, positions, indices, technique=0, radius=0.02, radius_per_vertex=None, lod=None):
        """lod = (line_split_passes, line_merge_passes, curve_merge_passes), as vkhrt_scene_apply_lod"""
        self.positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        self.indices = np.ascontiguousarray(indices, np.uint32).reshape(-1, 2)
        self.technique = int(technique)
        rpv = None
        if radius_per_vertex is not None:
            rpv = np.ascontiguousarray(radius_per_vertex, np.float32)
            assert rpv.shape[0] == self.positions.shape[0]
        self._rpv = rpv
        lod3 = np.asarray(lod if lod is not None else (0, 0, 0), np.uint32)
        self._h = lib().orc_scene_create_lod(self.positions.ctypes.data, self.positions.shape[0], self.indices.ctypes.data,
                                             self.indices.shape[0], rpv.ctypes.data if rpv is not None else None,
                                             float(radius), self.technique, lod3.ctypes.data)
        if not self._h:
            raise ValueError("oracle: bad topology")

    def set_environment(self, rgba):
        """RGBA32F equirectangular map [h, w, 4] (None removes it); frames select it with miss_mode=1"""
        if rgba is None:
            lib().orc_scene_set_environment(self._h, None, 0, 0)
            return
        e = np.ascontiguousarray(rgba, np.float32)
        assert e.ndim == 3 and e.shape[2] == 4
        lib().orc_scene_set_environment(self._h, e.ctypes.data, e.shape[1], e.shape[0])

    def set_material(self, albedo_factor=(1.0, 1.0, 1.0, 1.0), albedo_map=None):
        """Material::albedoFactor and an optional RGBA32F albedo map [h, w, 4] (frames use it with shade_mode = 2)"""
        f = np.ascontiguousarray(albedo_factor, np.float32)
        if albedo_map is None:
            lib().orc_scene_set_material(self._h, f.ctypes.data, None, 0, 0)
        else:
            m = np.ascontiguousarray(albedo_map, np.float32)
            lib().orc_scene_set_material(self._h, f.ctypes.data, m.ctypes.data, m.shape[1], m.shape[0])

    def set_meshes(self, first_segment):
        """multi-mesh scene: mesh m = segments [first_segment[m], first_segment[m + 1])"""
        fs = np.ascontiguousarray(first_segment, np.uint32).reshape(-1)
        lib().orc_scene_set_meshes(self._h, fs.ctypes.data if fs.size else None, fs.size)

    def set_mesh_material(self, mesh, albedo_factor=(1.0, 1.0, 1.0, 1.0), albedo_map=None):
        f = np.ascontiguousarray(albedo_factor, np.float32)
        if albedo_map is None:
            lib().orc_scene_set_mesh_material(self._h, int(mesh), f.ctypes.data, None, 0, 0)
        else:
            m = np.ascontiguousarray(albedo_map, np.float32)
            lib().orc_scene_set_mesh_material(self._h, int(mesh), f.ctypes.data, m.ctypes.data, m.shape[1], m.shape[0])

    def mesh_of_segment(self, seg):
        return int(lib().orc_scene_mesh_of_segment(self._h, int(seg)))

    def environment_miss(self, d):
        a = _f(d); o = (C.c_float * 3)()
        lib().orc_environment_miss(self._h, a[1], o)
        return np.array(list(o), np.float32)

    def lines(self):
        """the line list the primitives were generated from (after LOD): [n, 6] = start.xyz, end.xyz"""
        out = np.empty((int(lib().orc_scene_line_count(self._h)), 6), np.float32)
        lib().orc_scene_get_lines(self._h, out.ctypes.data)
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_scene_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def n_primitives(self):
        return int(lib().orc_scene_primitive_count(self._h))

    @property
    def n_leaves(self):
        """BVH leaves: leaf_split pieces per group (group = PHANTOM curve, LSS, or the 4-triangle DOTS strip of a segment)."""
        return int(lib().orc_scene_leaf_count(self._h))

    @property
    def leaf_split(self):
        return int(lib().orc_scene_leaf_split(self._h))

    def primitives(self):
        out = np.empty((self.n_primitives, FLOATS_PER_PRIM[self.technique]), np.float32)
        lib().orc_scene_get_primitives(self._h, out.ctypes.data)
        return out

    def aabbs(self):
        """leaf boxes (lo, hi), one per BVH leaf"""
        out = np.empty((self.n_leaves, 6), np.float32)
        lib().orc_scene_get_aabbs(self._h, out.ctypes.data)
        return out

    def bvh(self):
        n = self.n_leaves
        nodes = np.zeros(int(lib().orc_scene_node_count(self._h)), NODE_DTYPE)
        ids = np.zeros(n, np.uint32)
        morton = np.zeros(n, np.uint64)
        lohi = np.zeros(6, np.float32)
        lib().orc_scene_get_bvh(self._h, nodes.ctypes.data, ids.ctypes.data, morton.ctypes.data, lohi.ctypes.data)
        return nodes, ids, morton, lohi

    def render(self, frame, hits=True, rgba=True, pixel_subset=None, brute=False, stats=False, n_threads=0, n_out=None):
        if pixel_subset is not None:
            pixel_subset = np.ascontiguousarray(pixel_subset, np.uint64)
            n = pixel_subset.shape[0]
        elif n_out is not None:
            n = int(n_out)
        else:
            n = local_pixels(frame)
        h = np.zeros(n, HIT_DTYPE) if hits else None
        img = np.zeros((n, 4), np.uint8) if rgba else None
        st = TraceStats() if stats else None
        rc = lib().orc_render(self._h, C.byref(frame), h.ctypes.data if hits else None, img.ctypes.data if rgba else None,
                              pixel_subset.ctypes.data if pixel_subset is not None else None, n, int(brute),
                              C.byref(st) if stats else None, int(n_threads))
        assert rc == 0
        return h, img, (st.as_dict() if stats else None)

    def trace_rays(self, rays, brute=False, n_threads=0, any_hit=False):
        """any_hit: terminate on the first accepted hit in traversal order (shadow / occlusion rays)"""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        h = np.zeros(rays.shape[0], HIT_DTYPE)
        lib().orc_trace_rays(self._h, rays.ctypes.data, rays.shape[0], h.ctypes.data, int(brute), int(n_threads), int(any_hit))
        return h


def local_pixels(frame):
    T = frame.tile_size or 64
    stride = frame.tile_stride or 1
    if stride <= 1 or frame.row_major_output:
        return frame.width * frame.height
    tx, ty = (frame.width + T - 1) // T, (frame.height + T - 1) // T
    nt = tx * ty
    nl = (nt + stride - 1) // stride
    return nl * T * T


def prhi(ro, rd, curve, radius=0.02):
    _, pro = _f(ro); _, prd = _f(rd); _, pc = _f(np.asarray(curve).reshape(12))
    t, u = C.c_float(), C.c_float()
    n = (C.c_float * 3)()
    it = lib().orc_prhi(pro, prd, pc, radius, C.byref(t), C.byref(u), n)
    return float(t.value), float(u.value), np.array(list(n), np.float32), int(it)


def prhi_taper(ro, rd, curve, r0, r1):
    """Prhi with the radius running linearly from r0 (t = 0) to r1 (t = 1): cone.radius = r(t), cone.slant = r1 - r0"""
    _, pro = _f(ro); _, prd = _f(rd); _, pc = _f(np.asarray(curve).reshape(12))
    t, u = C.c_float(), C.c_float()
    n = (C.c_float * 3)()
    it = lib().orc_prhi_taper(pro, prd, pc, r0, r1, C.byref(t), C.byref(u), n)
    return float(t.value), float(u.value), np.array(list(n), np.float32), int(it)


def ray_cylinder(ro, rd, p0, p1, radius):
    a = [_f(x) for x in (ro, rd, p0, p1)]
    return bool(lib().orc_ray_cylinder(a[0][1], a[1][1], a[2][1], a[3][1], radius))


def lss(ro, rd, lss8):
    a = [_f(x) for x in (ro, rd, np.asarray(lss8).reshape(8))]
    t, u = C.c_float(), C.c_float(); n = (C.c_float * 3)()
    hit = lib().orc_lss(a[0][1], a[1][1], a[2][1], C.byref(t), C.byref(u), n)
    return bool(hit), float(t.value), float(u.value), np.array(list(n), np.float32)


def tri(ro, rd, tri9, parity=0):
    a = [_f(x) for x in (ro, rd, np.asarray(tri9).reshape(9))]
    t, u = C.c_float(), C.c_float(); n = (C.c_float * 3)()
    hit = lib().orc_tri(a[0][1], a[1][1], a[2][1], parity, C.byref(t), C.byref(u), n)
    return bool(hit), float(t.value), float(u.value), np.array(list(n), np.float32)


def raygen(view_inv, proj_inv, W, H, px, py, sample=0):
    a = [_f(np.asarray(x).reshape(16)) for x in (view_inv, proj_inv)]
    o = (C.c_float * 3)(); d = (C.c_float * 3)()
    lib().orc_raygen(a[0][1], a[1][1], W, H, px, py, sample, o, d)
    return np.array(list(o), np.float32), np.array(list(d), np.float32)


def curve_aabb(curve, radius=0.02):
    """GenerateAABBs for one curve (lo.xyz, hi.xyz)"""
    a = _f(np.asarray(curve).reshape(12)); o = (C.c_float * 6)()
    lib().orc_curve_aabb(a[1], radius, o)
    return np.array(list(o), np.float32)


def curve_point(curve, t):
    a = _f(np.asarray(curve).reshape(12)); o = (C.c_float * 3)()
    lib().orc_curve_point(a[1], t, o)
    return np.array(list(o), np.float32)


def curve_axis(curve, t):
    a = _f(np.asarray(curve).reshape(12)); o = (C.c_float * 3)()
    lib().orc_curve_axis(a[1], t, o)
    return np.array(list(o), np.float32)


def ao_direction(n, pixel, sample, index):
    a = _f(n); o = (C.c_float * 3)()
    lib().orc_ao_direction(a[1], pixel, sample, index, o)
    return np.array(list(o), np.float32)


def shade(n, prim=0, mode=0):
    a = _f(n); o = (C.c_float * 3)()
    lib().orc_shade(a[1], prim, mode, o)
    return np.array(list(o), np.float32)
