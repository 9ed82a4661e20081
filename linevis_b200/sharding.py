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


class FramesInFlight:
    """Two frames in flight: successive frames are rendered alternately by two contexts on two CUDA streams that share ONE scene, so
    the packet kernels of frame i + 1 (k_rtao_primary, k_tube_first) fill the SMs that the tail of frame i's persistent AO ray stream
    leaves idle, and frame i's k_rtao_reduce / k_tubes run beside the start of frame i + 1's stream.  Every frame is the same complete
    frame as with one context; only valid while frames do not depend on each other (no temporal accumulation: num_accumulated_frames
    = ambient_occlusion_iterations = 1, as in a benchmark replay or an animation), because each context keeps its own running means.

    contexts: two Context objects created on the two `streams` (torch.cuda.Stream) with identical settings; render(ctx, slot) enqueues
    one frame of that context (and, on several GPUs, its frame fence) -- it is called inside `with torch.cuda.stream(streams[slot])`."""

    def __init__(self, contexts, streams, render):
        assert len(contexts) == 2 and len(streams) == 2
        self.contexts, self.streams, self.render, self.i = contexts, streams, render, 0

    def step(self):
        j = self.i & 1
        self.i += 1
        with torch.cuda.stream(self.streams[j]):
            self.render(self.contexts[j], j)

    def fork(self, stream=None):
        """both render streams wait for everything enqueued on `stream` (default: the current stream) so far"""
        stream = stream or torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(stream)
        for s in self.streams:
            s.wait_event(ev)

    def join(self, stream=None):
        """`stream` (default: the current stream) waits for both render streams"""
        stream = stream or torch.cuda.current_stream()
        for s in self.streams:
            ev = torch.cuda.Event(); ev.record(s)
            stream.wait_event(ev)


# ------------------------------------------------------------------------------------------------ cost-balanced tile ownership
def balance_tiles_contiguous(costs, world, tile_pixels=64 * 64, spp=64, pixel_weight=1.5):
    """Owner rank per tile (Morton order, uint8): the Morton order is cut into `world` CONTIGUOUS runs of equal cost (prefix sums), so
    that a rank's tiles form a compact screen region -- its rays then touch about 1 / world of the scene instead of all of it
    (interleaved tiles make every rank pull the whole visible BVH through its L2 every frame, a cost that does not shrink with N)."""
    c = np.asarray(costs, np.float64) * float(spp) + float(pixel_weight) * float(tile_pixels)
    cum = np.cumsum(c)
    mid = cum - 0.5 * c                               # a tile belongs to the run its cost midpoint falls into
    owners = np.minimum((mid / (cum[-1] / world)).astype(np.int64), world - 1)
    return owners.astype(np.uint8)


def balance_tiles(costs, world, tile_pixels=64 * 64, spp=64, pixel_weight=1.5):
    """Owner rank per tile (Morton order, uint8) from the frame's cost map: cost of a tile = its AO rays (hit pixels x spp) + its pixels
    x `pixel_weight` (what the packet kernels of a pixel cost in AO-ray units: 0.8 ns against 0.53 ns on config 5).  Longest processing
    time first: tiles in order of decreasing cost, each to the rank with the smallest load so far (ties: the lowest rank; equal costs keep
    the Morton order) -- the same map on every rank for the same cost map.  With Morton round-robin the ranks' AO-ray counts differ by
    up to 16 % on config 5 (6.18 M .. 7.19 M of 52 M); balanced they differ by less than one tile."""
    costs = np.asarray(costs, np.float64) * float(spp) + float(pixel_weight) * float(tile_pixels)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world, np.float64)
    owners = np.zeros(costs.size, np.uint8)
    for i in order:
        r = int(np.argmin(load))
        owners[i] = r
        load[r] += costs[i]
    return owners


def rebalance(ctx, width, height, world, spp, tile=64, pixel_weight=1.5):
    """Every rank: cost map of the last frame (lv_get_tile_costs, summed over the ranks) -> balance_tiles -> lv_set_tile_owners.
    Call between two frames; the next frame is rendered with the new ownership (peer-frame assembly follows it by itself)."""
    costs = torch.from_numpy(ctx.tile_costs(width, height).astype(np.int64))
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        costs = costs.to(dev)
        dist.all_reduce(costs)
        costs = costs.cpu()
    owners = balance_tiles(costs.numpy(), world, tile * tile, spp, pixel_weight)
    ctx.set_tile_owners(width, height, owners)
    return owners


def tile_costs_single_gpu(ctx, scene, cam, tile, frame):
    """Cost map of the whole frame from ONE context (tools/shard_emul.py): render unsharded once, read the per-tile hit counts."""
    ctx.set_tile_shard(0, 1, tile)
    ctx.render_tubes(scene, cam, 0, out=frame, stats=False)
    return ctx.tile_costs(cam.width, cam.height)


# ------------------------------------------------------------------------------------------------ AO-sample-batch shards
def _device_floats(ptr, shape, device):
    """float32 tensor view of library-owned device memory (host memory when the library is the CPU emulation of the test-suite)"""
    if torch.device(device).type == "cuda":
        return torch.as_tensor(_RawCudaArray(ptr, shape), device=device)
    import ctypes
    n = int(np.prod(shape))
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_float * n).from_address(ptr))).view(*shape)


class SampleShards:
    """The second shard axis of the tube + RTAO frame (lv_sao_primary / lv_sao_trace / lv_sao_finish): pixels stay tile-sharded, the AO
    rays are split by SAMPLE -- rank r traces samples [r spp / N, (r + 1) spp / N) of every hit pixel of the frame, so every rank traces
    exactly the same number of rays through the same pixels (no tile imbalance).  Per frame: an all-gather of the ranks' hit lists
    (48 B per hit pixel) in front of the ray stream, an all-to-all of the per-sample results (4 B per ray) behind it; the owner sums its
    pixels' samples in sample order, so the frame is bit-identical to the tile-sharded and to the one-GPU frame.
    Needs spp % world == 0."""

    def __init__(self, ctx, rank, world, spp, device):
        if spp % world:
            raise ValueError("ambient_occlusion_samples_per_frame must be a multiple of the world size")
        self.ctx, self.rank, self.world, self.spp, self.spl, self.device = ctx, rank, world, spp, spp // world, device
        self._bufs = {}

    def _buf(self, name, n):
        b = self._bufs.get(name)
        if b is None or b.numel() < n:
            b = torch.empty(max(n, 1), dtype=torch.float32, device=self.device)
            self._bufs[name] = b
        return b[:n]

    def render(self, scene, cam, frame_number, out, stats=False):
        ctx, W = self.ctx, self.world
        ptr, n = ctx.sao_primary(scene, cam, frame_number)
        mine = _device_floats(ptr, (max(n, 1), 12), self.device)[:n]
        if W > 1:
            cnt = torch.tensor([n], dtype=torch.int64, device=self.device)
            cnts = torch.empty(W, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(cnts, cnt)
            counts = [int(c) for c in cnts.tolist()]
            total = sum(counts)
            # one padded all-gather (equal sizes: a single collective on NCCL, and what gloo supports), then the lists are packed in rank order
            cap = max(max(counts), 1)
            send = self._buf("send", cap * 12).view(cap, 12)
            send[:n] = mine
            padded = self._buf("padded", W * cap * 12).view(W, cap, 12)
            dist.all_gather_into_tensor(padded.view(-1), send.view(-1))
            hits = self._buf("hits", total * 12).view(total, 12)
            o = 0
            for r, c in enumerate(counts):
                hits[o:o + c] = padded[r, :c]
                o += c
        else:
            counts, total, hits = [n], n, mine
        occ = self._buf("occ", total * self.spl)
        ctx.sao_trace(scene, cam, frame_number, hits, total, self.rank * self.spl, self.spl, occ)
        if W > 1:
            parts = self._buf("parts", W * n * self.spl)
            dist.all_to_all_single(parts, occ, output_split_sizes=[n * self.spl] * W, input_split_sizes=[c * self.spl for c in counts])
        else:
            parts = occ
        return ctx.sao_finish(scene, cam, frame_number, parts, W, out, stats=stats)
