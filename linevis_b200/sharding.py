"""Image-tile sharding across the GPUs of one box (the reference itself is single-GPU).

The scene (segments + BVH) is replicated on every rank; the frame is cut into tile x tile tiles, enumerated in
Morton order, tile i -> rank i % world (balances dense / sparse image regions).  Each rank renders only its tiles
(lv_set_tile_shard); one all_gather of the packed tile blocks per frame puts the full framebuffer on every rank.
There is no other collective on the data path.

The tile enumeration here must match enumerate_tiles() in csrc/lv_api.cu (checked by a GPU test); the pure-torch
pack / unpack below is the host-side fallback used by the gloo CPU tests, on the GPU the CUDA kernels behind
lv_pack_owned_tiles / lv_unpack_tiles do the same job.
"""
import numpy as np
import torch
import torch.distributed as dist


def _part1by1(v):
    v = v & 0xFFFF
    v = (v | (v << 8)) & 0x00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F
    v = (v | (v << 2)) & 0x33333333
    v = (v | (v << 1)) & 0x55555555
    return v


def all_tiles(width, height, tile):
    """All (tile_x, tile_y) of a width x height frame in Morton order."""
    tx, ty = (width + tile - 1) // tile, (height + tile - 1) // tile
    xs, ys = np.meshgrid(np.arange(tx, dtype=np.uint32), np.arange(ty, dtype=np.uint32))
    xs, ys = xs.reshape(-1), ys.reshape(-1)
    code = _part1by1(xs) | (_part1by1(ys) << 1)
    order = np.argsort(code, kind="stable")
    return np.stack([xs[order], ys[order]], axis=1)


def owned_tiles(width, height, tile, rank, world):
    return all_tiles(width, height, tile)[rank::world]


def max_tiles_per_rank(width, height, tile, world):
    n = len(all_tiles(width, height, tile))
    return (n + world - 1) // world


def pack_tiles_torch(image, tiles, tile, n_slots):
    """image [H, W, C] -> [n_slots, tile*tile, C]; pixels outside the frame are zero."""
    H, W, C = image.shape
    out = torch.zeros((n_slots, tile, tile, C), dtype=image.dtype, device=image.device)
    for i, (tx, ty) in enumerate(tiles):
        x0, y0 = int(tx) * tile, int(ty) * tile
        w, h = min(tile, W - x0), min(tile, H - y0)
        out[i, :h, :w] = image[y0:y0 + h, x0:x0 + w]
    return out.reshape(n_slots, tile * tile, C)


def unpack_tiles_torch(packed, tiles, tile, image):
    H, W, C = image.shape
    p = packed.reshape(packed.shape[0], tile, tile, C)
    for i, (tx, ty) in enumerate(tiles):
        x0, y0 = int(tx) * tile, int(ty) * tile
        w, h = min(tile, W - x0), min(tile, H - y0)
        image[y0:y0 + h, x0:x0 + w] = p[i, :h, :w]
    return image


class FrameGather:
    """The single collective of a tile-sharded frame: all_gather of every rank's packed tile block.

    ctx: a linevis_b200.Context on this rank's GPU, or None for the host-side (gloo / CPU tensor) path."""

    def __init__(self, width, height, tile, rank, world, device, ctx=None, channels=4):
        self.W, self.H, self.tile, self.rank, self.world, self.ctx = width, height, tile, rank, world, ctx
        self.n_slots = max_tiles_per_rank(width, height, tile, world)
        self.mine = owned_tiles(width, height, tile, rank, world)
        self.packed = torch.zeros((self.n_slots, tile * tile, channels), dtype=torch.float32, device=device)
        # flat [world * n_slots, ...] receive buffer (the layout all_gather_into_tensor concatenates into); viewed per rank
        self._gathered_flat = torch.zeros((world * self.n_slots, tile * tile, channels), dtype=torch.float32, device=device)
        self.gathered = self._gathered_flat.view(world, self.n_slots, tile * tile, channels)

    def gather(self, frame, assemble_on=(0,)):
        """frame: [H, W, 4] tensor holding this rank's owned tiles.  After the call the ranks in `assemble_on` hold the
        complete frame in `frame`."""
        if self.ctx is not None:
            self.ctx.pack_owned_tiles(frame, self.W, self.H, self.packed)
        else:
            self.packed.copy_(pack_tiles_torch(frame, self.mine, self.tile, self.n_slots))
        if self.world > 1:
            dist.all_gather_into_tensor(self._gathered_flat, self.packed)
        else:
            self.gathered[0].copy_(self.packed)
        if assemble_on is None or self.rank in assemble_on:
            for r in range(self.world):
                if r == self.rank:
                    continue
                if self.ctx is not None:
                    self.ctx.unpack_tiles(self.gathered[r], r, self.world, self.W, self.H, frame)
                else:
                    unpack_tiles_torch(self.gathered[r], owned_tiles(self.W, self.H, self.tile, r, self.world), self.tile, frame)
        return frame


class _RawCudaArray:
    """Zero-copy view of a raw device address for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerFrame:
    """Peer-memory frame assembly (lv_frame_alloc / lv_ipc_*): ONE RGBA32F frame on `root`'s GPU that every rank's frame kernels
    store their owned tiles into directly over NVLink -- no pack / all_gather / unpack.  `fence()` is the frame fence: a
    one-element all_reduce on the current stream; after it the root's stream sees the complete frame.

    ptr: the address to pass as `out=` of Context.render_tubes / render_ppll on this rank."""

    def __init__(self, ctx, width, height, rank, world, device, root=0):
        self.ctx, self.W, self.H, self.rank, self.world, self.root = ctx, width, height, rank, world, root
        self.local = ctx.frame_alloc(width, height) if rank == root else None
        box = [ctx.ipc_export(self.local) if rank == root else None]
        if world > 1:
            dist.broadcast_object_list(box, src=root)
        self.ptr = self.local if rank == root else ctx.ipc_open(box[0])
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)
        self.device = device

    def fence(self):
        if self.world > 1:
            dist.all_reduce(self._flag)

    def tensor(self):
        """[H, W, 4] float32 view of the frame (root only)."""
        assert self.rank == self.root
        return torch.as_tensor(_RawCudaArray(self.local, (self.H, self.W, 4)), device=self.device)

    def tensor_rgba8(self):
        """[H, W] int32 view of the frame's first W*H*4 bytes (root only): the frame when the ranks render with b200_frame_format = rgba8."""
        assert self.rank == self.root
        return torch.as_tensor(_RawCudaArray(self.local, (self.H, self.W), "<i4"), device=self.device)

    def close(self):
        if self.ptr is None:
            return
        if self.rank == self.root:
            self.ctx.frame_free(self.local)
        else:
            self.ctx.ipc_close(self.ptr)
        self.ptr = self.local = None


# ---- multi-GPU baking of the object-space AO (lv_ao_set_vertex_range / lv_ao_factors) ----------------------------------------------
def bake_vertex_range(n_param, rank, world):
    """Slice [first, first + count) of the parametrization vertices rank `rank` bakes: contiguous, as equal as possible."""
    base, extra = divmod(int(n_param), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def exchange_baked_slices(factors, n_param, n_subdiv, rank, world):
    """factors: flat tensor [n_param * n_subdiv] (device view of lv_ao_factors, or a CPU tensor under gloo) in which this rank has baked
    its own vertex slice.  After the call every rank holds all slices: rank r's slice is broadcast from r (the slices differ in
    length, which rules out one all_gather_into_tensor; the payload is a few MB)."""
    if world == 1:
        return factors
    for r in range(world):
        first, count = bake_vertex_range(n_param, r, world)
        if count:
            dist.broadcast(factors[first * n_subdiv:(first + count) * n_subdiv], src=r)
    return factors
