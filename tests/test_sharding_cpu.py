"""Host-side logic of the multi-GPU path on CPU: tile ownership and the one all_gather per frame, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import linevis_b200 as lv
from linevis_b200 import scenes, sharding


def test_tiles_partition_the_frame():
    for (W, H, ts, world) in [(3840, 2160, 64, 8), (1920, 1080, 64, 4), (100, 70, 32, 3), (64, 64, 64, 2)]:
        allt = sharding.all_tiles(W, H, ts)
        assert len(allt) == ((W + ts - 1) // ts) * ((H + ts - 1) // ts)
        seen = set()
        counts = []
        for r in range(world):
            mine = sharding.owned_tiles(W, H, ts, r, world)
            counts.append(len(mine))
            for t in map(tuple, mine):
                assert t not in seen
                seen.add(t)
        assert len(seen) == len(allt) and max(counts) - min(counts) <= 1
        assert max(counts) == sharding.max_tiles_per_rank(W, H, ts, world)
    # Morton order: the first four tiles form a 2x2 block
    assert [tuple(t) for t in sharding.all_tiles(256, 256, 64)[:4]] == [(0, 0), (1, 0), (0, 1), (1, 1)]


def test_pack_unpack_roundtrip():
    img = torch.arange(70 * 100 * 4, dtype=torch.float32).reshape(70, 100, 4)
    out = torch.full_like(img, -1.0)
    for r in range(3):
        t = sharding.owned_tiles(100, 70, 32, r, 3)
        p = sharding.pack_tiles_torch(img, t, 32, sharding.max_tiles_per_rank(100, 70, 32, 3))
        sharding.unpack_tiles_torch(p, t, 32, out)
    assert torch.equal(out, img)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, tile, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import lvo
        o = lvo.Oracle("own")
        pos, attr, seg = scenes.helix_lines(12, 31)
        cam = lv.make_camera(W, H)
        full, _ = o.scene(pos, attr, seg, 0.01).render_tubes(cam, lvo.default_options(), scenes.standard_transfer_function())
        full = torch.from_numpy(full)
        # this rank "renders" only its tiles: everything else is left untouched (NaN) like lv_render_tubes does
        frame = torch.full_like(full, float("nan"))
        mine = sharding.owned_tiles(W, H, tile, rank, world)
        sharding.unpack_tiles_torch(sharding.pack_tiles_torch(full, mine, tile, len(mine)), mine, tile, frame)
        assert torch.isnan(frame).any()
        fg = sharding.FrameGather(W, H, tile, rank, world, torch.device("cpu"), ctx=None)
        fg.gather(frame, assemble_on=(0,))
        if rank == 0:
            q.put(bool(torch.equal(frame, full)))
        else:
            q.put(bool(torch.isnan(frame).any()))   # non-root ranks keep only their own tiles
        # units processed by all ranks: whole-job value = sum over ranks / max time (bench.py's reduction)
        t = torch.tensor([float(len(mine))], dtype=torch.float64)
        dist.all_reduce(t)
        assert int(t.item()) == len(sharding.all_tiles(W, H, tile))
    except Exception:
        q.put(False)
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_frame_gather_over_gloo():
    world, W, H, tile = 2, 100, 70, 32
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, tile, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert all(res)


def _bake_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import lvo
        o = lvo.Oracle("own")
        d = scenes.helix_polylines(5, 15)
        osc = o.scene(d["pos"], d["attr"], d["seg"], 0.012)
        osc.set_lines(d["tangent"], d["normal"])
        bw, sl = o.ao_parametrize(d["pos"], d["line_offsets"], 0.03)
        n_param, n_sub = len(sl), 6
        full, _ = osc.ao_bake_iteration(sl, 0, radius=0.2, n_subdiv=n_sub, spp=2)
        # this rank "bakes" only its vertex slice (the oracle stands in for the GPU), everything else stays zero
        first, count = sharding.bake_vertex_range(n_param, rank, world)
        mine = torch.zeros(n_param * n_sub)
        mine[first * n_sub:(first + count) * n_sub] = torch.from_numpy(full[first * n_sub:(first + count) * n_sub])
        sharding.exchange_baked_slices(mine, n_param, n_sub, rank, world)
        q.put((rank, bool(torch.equal(mine, torch.from_numpy(full)))))
    finally:
        dist.destroy_process_group()


def test_baked_slices_exchange_gloo_world3():
    """Multi-GPU baking, host logic: every rank bakes a contiguous vertex slice, the slices are broadcast, all ranks end with the full
    factor buffer (world size 3 so that the slices differ in length)."""
    world, port = 3, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bake_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True), (2, True)]
