"""Multi-GPU frame sharding: one process per GPU, BVH replicated, image tiles dealt round-robin.

The path has no exchange step (every ray is independent and read-only over the scene),
so the only communication is the final gather of the shards.  Rank r traces tiles r, r+world, r+2*world, ...
(VkhrtFrameDesc.tile_first / tile_stride).  Two ways to assemble the frame:

  "peer"   (default on GPUs) the gathering rank's full-frame buffers are mapped into every rank (CUDA IPC) and each
           rank's traversal kernel stores its hit records / pixels straight to their row-major position over NVLink
           (VkhrtFrameDesc.row_major_output): the kernel's own stores are the gather, overlapped with the traversal;
           a 4-byte NCCL all_reduce per frame is the completion signal.
  "gather" every rank writes a compact shard, all_gather (NCCL; gloo in the CPU tests) concatenates them rank-major,
           vkhrt_untile (CUDA) / untile_host (numpy mirror) restores row-major order.

The reference has no multi-GPU code at all (single graphics queue, source/vulkan_context.cpp:287-288).
"""
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

    def _setup_peer(self, n_full):
        torch, dist = self.torch, self.dist
        dev = self.device.index
        sizes = [n_full * 32] + ([n_full * 4] if self.want_rgba else [])
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
        self.hits_ptr = self._shared[0].ptr
        self.rgba_ptr = self._shared[1].ptr if self.want_rgba else None
        self._done = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.o_hits = torch.as_tensor(_DeviceView(self.hits_ptr, (n_full, 32)), device=self.device) if self.rank == self.gather_rank else None
        if self.want_rgba and self.rank == self.gather_rank:
            self.o_rgba = torch.as_tensor(_DeviceView(self.rgba_ptr, (n_full, 4)), device=self.device)

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
