"""Multi-GPU frame sharding: one process per GPU, BVH replicated, image tiles dealt round-robin.

The path has no exchange step (every ray is independent and read-only over the scene),
so the only communication is the final gather of the shards.  Rank r traces tiles r, r+world, r+2*world, ...
(VkhrtFrameDesc.tile_first / tile_stride).  Two ways to assemble the frame:

  "peer"   (default on GPUs) the gathering rank's full-frame buffers are mapped into every rank (CUDA IPC) and each
           rank's traversal kernel stores its hit records / pixels straight to their row-major position over NVLink
           (VkhrtFrameDesc.row_major_output): the kernel's own stores are the gather, overlapped with the traversal;
           a 4-byte NCCL all_reduce per frame is the completion signal.  There are TWO sets of frame buffers, used by
           alternate frames: the tensors render() returns stay valid while the next frame is being rendered (no rank
           can store into them before the frame after next), so the consumer needs no "I am done reading" signal.
  "gather" every rank writes a compact shard, all_gather (NCCL; gloo in the CPU tests) concatenates them rank-major,
           vkhrt_untile (CUDA) / untile_host (numpy mirror) restores row-major order.

The reference has no multi-GPU code at all (single graphics queue, source/vulkan_context.cpp:287-288).
"""
import os

import numpy as np

from . import api as _api


class TileSharding:
    """Host-side description of the round-robin tile layout (mirrors `resolve()` in csrc/trace.cu)."""

    def __init__(self, width, height, world, tile=64):
        if tile % 8 or tile <= 0:
            raise ValueError("tile size must be a positive multiple of 8")
        self.width, self.height, self.world, self.tile = int(width), int(height), int(world), int(tile)
        self.tiles_x = (self.width + tile - 1) // tile
        self.tiles_y = (self.height + tile - 1) // tile
        self.n_tiles = self.tiles_x * self.tiles_y
        self.n_local_tiles = (self.n_tiles + self.world - 1) // self.world   # same on every rank: shards gather evenly
        self.shard_pixels = self.n_local_tiles * tile * tile if world > 1 else self.width * self.height

    @staticmethod
    def balanced_tile(width, world, preferred=64):
        """A tile size (multiple of 8, near `preferred`) whose tile count per image row is coprime with `world`.
        Tiles are dealt round-robin over the row-major tile index; when the row length divides by `world` (e.g. 3840 / 64 = 60
        tiles over 4 ranks) every rank gets VERTICAL STRIPES and the load follows the horizontal hair density; with a coprime row
        length the pattern shifts from row to row (diagonal), which evens it out."""
        from math import gcd
        if world <= 1:
            return preferred
        for t in sorted(range(32, 129, 8), key=lambda t: abs(t - preferred)):
            if gcd(-(-width // t) % world or world, world) == 1:
                return t
        return preferred

    def rank_of_tile(self, tile_index):
        return tile_index % self.world

    def tiles_of_rank(self, rank):
        return list(range(rank, self.n_tiles, self.world))

    def frame_kwargs(self, rank):
        """tile_* fields of make_frame() for this rank."""
        return dict(tile_size=self.tile, tile_first=rank if self.world > 1 else 0, tile_stride=self.world if self.world > 1 else 0)

    def gather_index(self):
        """int64[H*W]: position in the rank-major gathered buffer of every row-major pixel."""
        T = self.tile
        py, px = np.divmod(np.arange(self.width * self.height, dtype=np.int64), self.width)
        tile = (py // T) * self.tiles_x + px // T
        rank, local = tile % self.world, tile // self.world
        return rank * (self.n_local_tiles * T * T) + local * T * T + (py % T) * T + (px % T)

    def untile_host(self, gathered):
        """numpy mirror of vkhrt_untile: gathered[world * shard_pixels, ...] -> row-major [H*W, ...]."""
        gathered = np.asarray(gathered)
        if self.world == 1:
            return gathered[: self.width * self.height]
        return gathered[self.gather_index()]


class _DeviceView:
    """Minimal __cuda_array_interface__ wrapper so torch can view a raw device pointer (torch.as_tensor(view))."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class ShardedRenderer:
    """One rank's view of a frame rendered by `world` GPUs (torch.distributed process group already initialised)."""

    def __init__(self, scene, width, height, group=None, tile=64, spp=1, want_rgba=False, device=None, mode="peer", gather_rank=0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.scene = scene
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.layout = TileSharding(width, height, self.world, tile)
        self.spp, self.want_rgba = spp, want_rgba
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.gather_rank = gather_rank
        self.mode = mode if self.world > 1 else "single"
        self._shared = []
        self.o_rgba = None
        n_full = width * height
        if self.mode == "peer":
            try:
                self._setup_peer(n_full)
            except Exception as e:                  # IPC not permitted on this box: fall back to the NCCL gather
                self.mode = "gather"
                self.peer_error = repr(e)
        if self.mode == "peer":
            return
        n = self.layout.shard_pixels
        self.d_hits = torch.empty((n, 32), dtype=torch.uint8, device=self.device)
        self.d_rgba = torch.empty((n, 4), dtype=torch.uint8, device=self.device) if want_rgba else None
        if self.world > 1:
            self.g_hits = torch.empty((self.world * n, 32), dtype=torch.uint8, device=self.device)
            self.o_hits = torch.empty((n_full, 32), dtype=torch.uint8, device=self.device)
            if want_rgba:
                self.g_rgba = torch.empty((self.world * n, 4), dtype=torch.uint8, device=self.device)
                self.o_rgba = torch.empty((n_full, 4), dtype=torch.uint8, device=self.device)
        else:
            self.o_hits, self.o_rgba = self.d_hits, self.d_rgba
        self.hits_ptr = self.d_hits.data_ptr()
        self.rgba_ptr = self.d_rgba.data_ptr() if want_rgba else None

    N_BUFFERS = 2      # peer mode: frame k lives in buffer set k % 2

    def _setup_peer(self, n_full):
        torch, dist = self.torch, self.dist
        dev = self.device.index
        per_set = [n_full * 32] + ([n_full * 4] if self.want_rgba else [])
        sizes = per_set * self.N_BUFFERS
        ok = 1
        handles = [None]
        if self.rank == self.gather_rank:
            try:
                self._shared = [_api.SharedBuffer.create(b, dev) for b in sizes]
                handles = [[sb.handle for sb in self._shared]]
            except Exception:
                ok = 0
        dist.broadcast_object_list(handles, src=self.gather_rank, group=self.group)
        if self.rank != self.gather_rank:
            try:
                if handles[0] is None:
                    raise RuntimeError("exporter failed")
                self._shared = [_api.SharedBuffer.open(h, dev) for h in handles[0]]
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self.close()
            raise RuntimeError("CUDA IPC mapping of the gathering rank's frame buffer failed on at least one rank")
        k = len(per_set)
        self._frame_index = 0
        self._hits_ptrs = [self._shared[b * k].ptr for b in range(self.N_BUFFERS)]
        self._rgba_ptrs = [self._shared[b * k + 1].ptr if self.want_rgba else None for b in range(self.N_BUFFERS)]
        self._done = torch.zeros(1, dtype=torch.int32, device=self.device)
        mine = self.rank == self.gather_rank
        self._o_hits = [torch.as_tensor(_DeviceView(q, (n_full, 32)), device=self.device) if mine else None for q in self._hits_ptrs]
        self._o_rgba = [torch.as_tensor(_DeviceView(q, (n_full, 4)), device=self.device) if (mine and self.want_rgba) else None for q in self._rgba_ptrs]
        self.hits_ptr, self.rgba_ptr = self._hits_ptrs[0], self._rgba_ptrs[0]
        self.o_hits, self.o_rgba = self._o_hits[0], self._o_rgba[0]

    def close(self):
        for sb in self._shared:
            sb.close()
        self._shared = []

    def make_frame(self, view_inv, proj_inv, stream, **kw):
        self._full = _api.make_frame(view_inv, proj_inv, self.layout.width, self.layout.height, tile_size=self.layout.tile)
        return _api.make_frame(view_inv, proj_inv, self.layout.width, self.layout.height, spp=self.spp,
                               output_memory=_api.MEM_DEVICE, stream=stream, row_major_output=1 if self.mode == "peer" else 0,
                               **self.layout.frame_kwargs(self.rank), **kw)

    def render(self, frame, stream):
        """Trace this rank's tiles and assemble the frame on the gathering rank.  Everything is enqueued on `stream`
        (a raw cudaStream_t that must be torch's current stream so the NCCL call orders after the kernel)."""
        if self.mode == "peer":
            # Frame k goes to buffer set k % 2.  The all_reduce of frame k completes on a rank only after EVERY rank has enqueued
            # its own (after its frame-k kernel), and a rank's frame-(k+1) kernel is ordered after its frame-k all_reduce; so
            # when any rank starts storing frame k+2 into this set again, the gathering rank's stream is past frame k+1's
            # all_reduce, i.e. past everything it enqueued to read frame k.  The returned tensors are therefore valid until the
            # call after next (consume them on `stream`, or copy them, before rendering two more frames).
            b = self._frame_index % self.N_BUFFERS
            self._frame_index += 1
            self.hits_ptr, self.rgba_ptr = self._hits_ptrs[b], self._rgba_ptrs[b]
            self.o_hits, self.o_rgba = self._o_hits[b], self._o_rgba[b]
            self.scene.render_into(frame, self.hits_ptr, self.rgba_ptr)
            self.dist.all_reduce(self._done, group=self.group)       # stream-ordered "every shard has landed" signal
            return self.o_hits, self.o_rgba
        self.scene.render_into(frame, self.d_hits.data_ptr(), self.d_rgba.data_ptr() if self.want_rgba else None)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.g_hits, self.d_hits, group=self.group)
            _api.untile(self._full, self.world, self.g_hits.data_ptr(), self.o_hits.data_ptr(), 32, stream)
            if self.want_rgba:
                self.dist.all_gather_into_tensor(self.g_rgba, self.d_rgba, group=self.group)
                _api.untile(self._full, self.world, self.g_rgba.data_ptr(), self.o_rgba.data_ptr(), 4, stream)
        return self.o_hits, self.o_rgba


class SharedHostFrame:
    """ONE page-locked host frame that every rank's GPU stores its hit records into (the N > 1 end-to-end path).

    The gathering rank creates a POSIX shared-memory segment, every rank maps it and registers it with CUDA
    (cudaHostRegister, portable + mapped).  vkhrt_render with output_memory = HOST, row_major_output = 1 then stores
    each rank's records at their row-major position over that rank's own PCIe link (zero-copy / line-wise delivery,
    DESIGN.md §6): after a barrier the gathering rank holds the assembled frame in host memory — no staging, no copy.
    """

    def __init__(self, n_records, group=None, gather_rank=0):
        import torch
        import torch.distributed as dist
        import _posixshmem      # shm_open / shm_unlink as multiprocessing.shared_memory uses them, without its resource tracker: the
        import mmap             # tracker of Python < 3.13 also adopts ATTACHED segments and unlinks (or double-unregisters) them at exit
        import secrets
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nbytes = int(n_records) * 32
        self.owner = self.rank == gather_rank
        self._unlink = _posixshmem.shm_unlink
        name = [None]
        if self.owner:
            self.name = "/vkhrt_frame_%d_%s" % (os.getpid(), secrets.token_hex(4))
            fd = _posixshmem.shm_open(self.name, os.O_CREAT | os.O_EXCL | os.O_RDWR, mode=0o600)
            try:
                os.ftruncate(fd, max(self.nbytes, 1))
                self.map = mmap.mmap(fd, max(self.nbytes, 1))
            finally:
                os.close(fd)
            name = [self.name]
        if self.world > 1:
            dist.broadcast_object_list(name, src=gather_rank, group=group)
        if not self.owner:
            self.name = name[0]
            fd = _posixshmem.shm_open(self.name, os.O_RDWR, mode=0o600)
            try:
                self.map = mmap.mmap(fd, max(self.nbytes, 1))
            finally:
                os.close(fd)
        self.array = np.frombuffer(self.map, dtype=np.uint8, count=self.nbytes)
        self.ptr = self.array.ctypes.data
        rc = torch.cuda.cudart().cudaHostRegister(self.ptr, self.nbytes, 1 | 2)      # cudaHostRegisterPortable | Mapped
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister of the shared host frame failed: {rc}")
        self.registered = True

    def hits(self):
        return self.array.view(_api.HIT_DTYPE)

    def close(self):
        if getattr(self, "registered", False):
            self.torch.cuda.cudart().cudaHostUnregister(self.ptr)
            self.registered = False
        self.array = None
        try:
            self.map.close()
        except (BufferError, ValueError):
            pass                    # a view handed out by hits() is still alive: the mapping goes with the process
        if self.owner and getattr(self, "name", None):
            try:
                self._unlink(self.name)
            except OSError:
                pass
            self.name = None
