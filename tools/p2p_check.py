"""Multi-GPU check of the peer-memory frame assembly (run under torchrun, >= 2 ranks): the frame assembled by NVLink peer stores
must equal the all_gather-assembled frame bit for bit; prints the frame times of both modes (CUDA events, max over ranks)."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench, linevis_b200 as lv
from linevis_b200.sharding import FrameGather, PeerFrame

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config5")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wl = bench.WORKLOADS[args.workload]
W, H = wl["W"], wl["H"]
pos, attr, seg = bench.generate(wl["gen"], dev)
ctx = lv.Context(local, torch.cuda.current_stream().cuda_stream)
ctx.set_transfer_function(lv.scenes.standard_transfer_function())
ctx.set_new_settings({"ambient_occlusion_strength": 1.0, "ambient_occlusion_samples_per_frame": wl["ao_spp"], "ambient_occlusion_iterations": 1,
                      "num_samples_per_frame": 1, "num_accumulated_frames": 1})
ctx.set_tile_shard(rank, world, 64)
d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (pos, attr, seg.view(np.int32))]
sc = ctx.create_scene(d[0], d[1], d[2], lv.scenes.LINE_WIDTH)
cam = lv.make_camera(W, H)
frame = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
fg = FrameGather(W, H, 64, rank, world, dev, ctx=ctx)
pf = PeerFrame(ctx, W, H, rank, world, dev)


def step_gather():
    ctx.render_tubes(sc, cam, 0, out=frame, stats=False)
    fg.gather(frame, assemble_on=(0,))


def step_p2p():
    ctx.render_tubes(sc, cam, 0, out=pf.ptr, stats=False)
    pf.fence()


def timed(fn):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def step_render_only():
    ctx.render_tubes(sc, cam, 0, out=frame, stats=False)


def step_gather_only():
    fg.gather(frame, assemble_on=(0,))


t_g = timed(step_gather)
t_p = timed(step_p2p)
t_r = timed(step_render_only)
t_go = timed(step_gather_only)
t_f = timed(pf.fence)
# single-GPU reference of the same frame on rank 0 (whole frame, no sharding) for the speed-up
st = ctx.render_tubes(sc, cam, 0, out=frame, stats=True)[1]
mine = torch.tensor([st["ms_rtao_rays"], st["ms_total"], st["rays_ao"] / 1e6], dtype=torch.float64, device=dev)
allr = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allr, mine)
ctx.set_tile_shard(0, 1, 64)
full = [ctx.render_tubes(sc, cam, 0, out=frame, stats=True)[1]["ms_total"] for _ in range(3)][-1]
ctx.set_tile_shard(rank, world, 64)
for _ in range(2):
    step_gather()
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    same = torch.equal(pf.tensor().view(torch.int32), frame.view(torch.int32))
    print("world %d: all_gather frame %.3f ms, peer-store frame %.3f ms, frames bit-identical: %s, non-background pixels %d" %
          (world, t_g, t_p, same, int((frame[..., 0] < 0.999).sum().item())), flush=True)
    print("   render only (max over ranks) %.3f ms, pack+all_gather+unpack alone %.3f ms, fence alone %.3f ms, whole frame on one GPU %.3f ms "
          "-> speed-up %.2fx (all_gather) / %.2fx (peer stores)" % (t_r, t_go, t_f, full, full / t_g, full / t_p), flush=True)
    print("   per rank [k_rtao_rays ms, frame ms, M AO rays]: " + "  ".join("[%.2f %.2f %.2f]" % tuple(t.tolist()) for t in allr), flush=True)
dist.barrier()
pf.close()
dist.destroy_process_group()
