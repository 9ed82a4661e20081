#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native LineVis hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one frame of the tube + RTAO path (lv_render_tubes: RTAO pass + tube ray-gen) on BASELINE.json's
headline configuration ("10 M curl-noise streamline segments, 3840x2160, tubes + 64-spp RTAO"), which fits one GPU.
`value` = Mrays/s with everything resident in HBM; `e2e` = the same metric through the C ABI with HOST buffers
(camera in, framebuffer read back to pinned host memory inside the timed region).  The PPLL path ("100 k-segment
helix, 1920x1080, PPLL OIT") is measured beside it and reported under "ppll" in the same JSON line.

--impl reference times the reference's own CPU path for the same workload: the oracle drivers on the reference's
madmann91/bvh library (oracle/_ref, built from /root/reference) or, if that is absent, the oracle port.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (generator kwargs, frame, settings)
    "config5": dict(desc="10 M curl-noise streamline segments (20000 lines x 501 pts, seed 3003), 3840x2160, tubes + 64-spp RTAO",
                    gen=("curl", dict(n_lines=20000, n_points=501, seed=3003)), W=3840, H=2160, ao_spp=64),
    "config3": dict(desc="1 M random line segments (seed 2002), 1920x1080, tubes + 16-spp RTAO",
                    gen=("random", dict(n_seg=1_000_000, seed=2002)), W=1920, H=1080, ao_spp=16),
    "small": dict(desc="100 k-segment helix, 640x360, tubes + 8-spp RTAO (debug size)",
                  gen=("helix", dict()), W=640, H=360, ao_spp=8),
}
PPLL_WORKLOADS = {
    "config2": dict(desc="100 k-segment synthetic helix (seed 1001), 1920x1080, PPLL OIT, MAX_NUM_FRAGS 100",
                    gen=("helix", dict()), W=1920, H=1080, max_frags=100),
    # fragment budget: expectedAvgDepthComplexity 24 instead of the reference's 20 for <= 1 M segments (this frame holds 19.2 fragments
    # per pixel on average; a tile shard's share varies by a few per cent and nothing may be dropped inside the timed region)
    "config4": dict(desc="1 M random segments (seed 2002), 3840x2160, PPLL OIT, MAX_NUM_FRAGS 256, fragment budget 24 x pixels",
                    gen=("random", dict(n_seg=1_000_000, seed=2002)), W=3840, H=2160, max_frags=256, avg_depth=24),
}


def generate(gen, device=None):
    from linevis_b200 import scenes
    kind, kw = gen
    if kind == "curl":
        return scenes.curl_noise_streamlines(device=device, **kw)
    if kind == "random":
        return scenes.random_segments(**kw)
    return scenes.helix_lines(**kw)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/traffic.json);
    only valid for the workload the capture was taken on (config5 for k_rtao_rays_q)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if workload != "config5" or not os.path.exists(p):
        return None
    return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, wl, ppll_wl):
    """The reference's own CPU path for this workload, on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import linevis_b200 as lv
    from oracle import lvo
    try:
        o = lvo.Oracle("ref"); kind = "reference"
    except (FileNotFoundError, OSError):
        o = lvo.Oracle("own"); kind = "port"
    o.set_num_threads()          # torchrun exports OMP_NUM_THREADS=1 for nproc > 1: size the pool explicitly (all host cores)
    pos, attr, seg = generate(wl["gen"])
    t0 = time.time()
    sc = o.scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    build_s = time.time() - t0
    # bounded sample of the same workload: the central crop of the full-resolution frame, rendered at the full frame's
    # ray density (same camera, same spp); Mrays/s is intensive, the measured rate is reported, never an extrapolated time.
    sw, sh = args.ref_sample
    cam = lv.make_camera(wl["W"], wl["H"])
    sub = lv.make_camera(sw, sh, fov_y=2.0 * math.atan(0.5 * sh / wl["H"]))
    opts = lvo.default_options(ao_strength=1.0, ao_spp=wl["ao_spp"], ao_jitter_primary=1, ao_use_distance=1)
    tf = lv.scenes.standard_transfer_function()
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        ao, s1 = sc.render_rtao(sub, opts, 0)
        img, s2 = sc.render_tubes(sub, opts, tf, ao_tex=ao)
        dt = time.time() - t0
        if i >= args.warmup:
            times.append(dt)
            rays = s1["rays_primary"] + s1["rays_ao"] + s2["rays"]
    ms = 1e3 * float(np.mean(times))
    value = rays / (ms * 1e-3) / 1e6
    T = (s1["T"] + s2["T"]) / max(rays, 1); I = (s1["I"] + s2["I"]) / max(rays, 1)
    line = {
        "impl": "reference", "metric": "Mrays/s (tube+RTAO)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": wl["desc"], "sample": "%dx%d centre crop of the %dx%d frame" % (sw, sh, wl["W"], wl["H"])},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": o.num_threads(), "kind": kind,
                         "sample": "%dx%d centre crop, %d spp RTAO, %d rays/step; BVH %s built in %.1f s (not timed); T/ray %.1f I/ray %.1f"
                                   % (sw, sh, wl["ao_spp"], rays, o.lib.lvo_backend_name().decode(), build_s, T, I)},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------- CPU baseline leg
def cpu_baseline(wl, pos, attr, seg, sample, budget_s=25.0):
    import linevis_b200 as lv
    from oracle import lvo
    try:
        o = lvo.Oracle("ref"); kind = "reference"
    except (FileNotFoundError, OSError):
        o = lvo.Oracle("own"); kind = "port"
    o.set_num_threads()
    t0 = time.time()
    sc = o.scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    build_s = time.time() - t0
    sw, sh = sample
    sub = lv.make_camera(sw, sh, fov_y=2.0 * math.atan(0.5 * sh / wl["H"]))
    opts = lvo.default_options(ao_strength=1.0, ao_spp=wl["ao_spp"], ao_jitter_primary=1, ao_use_distance=1)
    tf = lv.scenes.standard_transfer_function()
    t0 = time.time()
    ao, s1 = sc.render_rtao(sub, opts, 0)
    img, s2 = sc.render_tubes(sub, opts, tf, ao_tex=ao)
    dt = time.time() - t0
    rays = s1["rays_primary"] + s1["rays_ao"] + s2["rays"]
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": o.num_threads(), "kind": kind,
            "sample": "%dx%d centre crop of the frame, %d spp, %d rays in %.1f s; CPU BVH (%s) build %.1f s not timed; T/ray %.1f, I/ray %.1f"
                      % (sw, sh, wl["ao_spp"], rays, dt, o.lib.lvo_backend_name().decode(), build_s,
                         (s1["T"] + s2["T"]) / max(rays, 1), (s1["I"] + s2["I"]) / max(rays, 1))}


def cpu_baseline_ppll(pw, pos, attr, seg, sample):
    import linevis_b200 as lv
    from oracle import lvo
    try:
        o = lvo.Oracle("ref"); kind = "reference"
    except (FileNotFoundError, OSError):
        o = lvo.Oracle("own"); kind = "port"
    o.set_num_threads()
    sc = o.scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    sw, sh = sample
    sub = lv.make_camera(sw, sh, fov_y=2.0 * math.atan(0.5 * sh / pw["H"]))
    tf = lv.scenes.standard_transfer_function(opacity=(0.1, 0.6))
    opts = lvo.default_options()
    g = sc.ppll_gather(sub, opts, tf)
    reps, t0 = 0, time.time()
    while reps < 3 or time.time() - t0 < 3.0:          # the resolve of one crop takes milliseconds: repeat it for ~3 s
        img, st = lvo.ppll_resolve(o, sub, opts, g["heads"], g["nodes"], pw["max_frags"], 0, canonical=False)
        reps += 1
    dt = (time.time() - t0) / reps
    return {"value": st["frags_sorted"] / dt / 1e6, "unit": "Mfrags/s sorted", "cores": o.num_threads(), "kind": kind,
            "sample": "%dx%d centre crop, %d fragments resolved (frontToBackPQ) in %.4f s, mean of %d repetitions" % (sw, sh, st["frags_sorted"], dt, reps)}


# ------------------------------------------------------------------------------------------------- our arm
_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's "NCCL version ..." banner at NCCL_DEBUG=VERSION/WARN):
    file descriptor 1 is pointed at stderr for the rest of the process and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config5", choices=list(WORKLOADS))
    ap.add_argument("--ppll-workload", default="config2,config4",
                    help="comma-separated PPLL workloads measured beside the tube path (%s) or 'none'; the first is reported under "
                         "\"ppll\", further ones under \"ppll_<name>\"" % ", ".join(PPLL_WORKLOADS))
    ap.add_argument("--ref-sample", type=int, nargs=2, default=[1920, 1080],
                    help="centre crop (pixels) of the frame the CPU legs render: ~90 M rays, about 10 s per step on 16 host cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--assemble", default="peer", choices=["peer", "allgather"],
                    help="N > 1: how rank 0 gets the whole frame -- 'peer': every rank's kernels store their tiles straight into rank 0's "
                         "frame over NVLink (lv_frame_alloc / lv_ipc_*), one 1-element all_reduce as frame fence; 'allgather': pack + NCCL "
                         "all_gather + unpack")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="extra lv_set_option settings for A/B runs (e.g. b200_ao_qnodes=true, b200_ppll_reg_sort=true); recorded in config.options")
    args = ap.parse_args()
    extra_opts = dict(o.split("=", 1) for o in args.opt)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]
    ppll_names = [n for n in args.ppll_workload.split(",") if n and n != "none"]
    for n in ppll_names:
        if n not in PPLL_WORKLOADS:
            ap.error("unknown PPLL workload %r" % n)
    pw = PPLL_WORKLOADS[ppll_names[0]] if ppll_names else None

    if args.impl == "reference":
        run_reference(args, wl, pw)
        return

    import torch
    import torch.distributed as dist
    import linevis_b200 as lv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hbm_peak, peak_src = peaks()

    W, H = wl["W"], wl["H"]
    pos, attr, seg = generate(wl["gen"], dev)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = lv.Context(local, stream)
    ctx.set_transfer_function(lv.scenes.standard_transfer_function())
    ctx.set_new_settings({
        "depth_cue_strength": 0.0, "ambient_occlusion_strength": 1.0, "ambient_occlusion_gamma": 1.0,
        "ambient_occlusion_samples_per_frame": wl["ao_spp"], "ambient_occlusion_iterations": 1, "ambient_occlusion_radius": 0.1,
        "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True,
        "num_samples_per_frame": 1, "num_accumulated_frames": 1, "use_deterministic_sampling": False})
    ctx.set_new_settings(extra_opts)
    tile = 64
    if world > 1:
        ctx.set_tile_shard(rank, world, tile)
    t0 = time.time()
    d_pos, d_attr, d_seg = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (pos, attr, seg.view(np.int32)))
    scene = ctx.create_scene(d_pos, d_attr, d_seg, lv.scenes.LINE_WIDTH)
    torch.cuda.synchronize()
    upload_build_s = time.time() - t0
    first_build_ms = scene.info()["build_ms"]      # the first build of a process also pays for CUDA module loading
    scene.close()
    scene = ctx.create_scene(d_pos, d_attr, d_seg, lv.scenes.LINE_WIDTH)
    torch.cuda.synchronize()
    info = scene.info()
    scene_bytes = info["n_seg"] * 36 + info["n_nodes"] * 64
    cam = lv.make_camera(W, H)

    frame = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    n_own = len(ctx.owned_tiles(W, H))
    from linevis_b200.sharding import FrameGather, PeerFrame
    peer = world > 1 and args.assemble == "peer"
    fg = FrameGather(W, H, tile, rank, world, dev, ctx=ctx) if world > 1 else None     # also the e2e leg's device-side collective
    pf = PeerFrame(ctx, W, H, rank, world, dev) if peer else None

    def step(stats):
        if pf is not None:
            # every rank's frame kernels store their tiles straight into rank 0's frame (NVLink peer stores); the fence is the frame's only collective
            out, st = ctx.render_tubes(scene, cam, 0, out=pf.ptr, stats=stats)
            pf.fence()
            return st
        out, st = ctx.render_tubes(scene, cam, 0, out=frame, stats=stats)
        if fg is not None:
            # the single collective of the frame: every rank's packed tile block -> all ranks; rank 0 assembles the frame
            fg.gather(frame, assemble_on=(0,))
        return st

    # warm-up (also yields the per-frame ray / T / I counts: the frame is deterministic)
    st = None
    for _ in range(args.warmup):
        st = step(True)
    rays = st["rays_primary"] + st["rays_ao"]
    counts = torch.tensor([rays, st["traversal_steps"], st["intersections"], st["rays_primary"], st["rays_ao"],
                           st["ao_traversal_steps"], st["ao_intersections"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    tot_rays, tot_T, tot_I, tot_rp, tot_ra, ao_T, ao_I = [float(x) for x in counts.tolist()]
    ms_rtao_rays_warm = st["ms_rtao_rays"]

    # ---- timed region: exactly K steps, barrier + synchronize on both sides, CUDA events, max over ranks
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(False)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) / args.steps
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    clocks = sampler.stop() if sampler else None

    # ---- dominant kernel (k_rtao_rays_q, the AO ray stream) live timing for the roofline: CUDA events around that kernel inside the library
    kt = []
    for _ in range(max(3, args.steps)):
        s2 = ctx.render_tubes(scene, cam, 0, out=frame, stats=True)[1]
        kt.append(s2["ms_rtao_rays"])
    k_ms = float(np.mean(kt))
    my_ao_bytes = 64 * st["ao_traversal_steps"] + 32 * st["ao_intersections"] + 4 * st["rays_ao"]
    achieved = my_ao_bytes / (k_ms * 1e-3) / 1e9

    # ---- end to end through the C ABI with HOST buffers (camera struct in, RGBA32F frame out to pinned host memory)
    host_frame = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
    host_np = host_frame.numpy()

    def step_e2e():
        if pf is not None:
            # every rank renders into rank 0's frame (peer stores), fence, rank 0 reads the WHOLE frame back to its pinned host
            # buffer; the second fence keeps the next frame's stores off the buffer until that copy is done
            ctx.render_tubes(scene, cam, 0, out=pf.ptr, stats=False)
            pf.fence()
            if rank == 0:
                host_frame.copy_(pf.tensor(), non_blocking=True)
            pf.fence()
            torch.cuda.current_stream().synchronize()
            return
        ctx.render_tubes(scene, cam, 0, out=host_np, stats=False)   # D2H inside, synchronises
        if fg is not None:
            fg.gather(frame, assemble_on=())                        # the collective stays in the e2e step as well
    for _ in range(2):
        step_e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    tms = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_ms = float(tms.item())
    # ---- the same step with the frame delivered in the reference's own sceneTexture format (RGBA8 UNORM, lv_frame_to_rgba8): a quarter
    # of the read-back bytes.  Reported beside `e2e` (which stays the RGBA32F delivery), single GPU only; never fatal.
    e2e8 = None
    if world == 1:
        try:
            host8 = torch.zeros((H, W), dtype=torch.int32).pin_memory()
            host8_np = host8.numpy().view(np.uint32)

            def step_e2e8():
                ctx.render_tubes(scene, cam, 0, out=frame, stats=False)
                ctx.frame_to_rgba8(frame, W, H, out=host8_np)       # conversion kernel + D2H inside, synchronises
            for _ in range(2):
                step_e2e8()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_e2e8()
            torch.cuda.synchronize()
            ms8 = (time.perf_counter() - t0) * 1e3 / args.steps
            e2e8 = {"value": tot_rays / (ms8 * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms8, "h2d_bytes_per_step": ctypes.sizeof(lv.LvCamera),
                    "d2h_bytes_per_step": W * H * 4, "nonzero_pixels": int(np.count_nonzero(host8_np)),
                    "note": "lv_render_tubes into a device frame + lv_frame_to_rgba8 into pinned host memory (RGBA8 UNORM, the reference's sceneTexture format)"}
        except Exception as e:   # noqa: BLE001 -- an optional extra measurement must not cost the bench line
            e2e8 = {"error": "%s: %s" % (type(e).__name__, e)}
    # bytes read back per step: the whole frame on rank 0 (single GPU, or peer assembly), else every rank's own tiles (rank 0's share is reported)
    d2h = (n_own * tile * tile if (world > 1 and pf is None) else W * H) * 16

    # ---- PPLL path (second half of the metric), rank-local, whole frame on one GPU unless sharded
    def measure_ppll(pw, with_cpu):
        ppos, pattr, pseg = generate(pw["gen"]) if pw["gen"] != wl["gen"] else (pos, attr, seg)
        pctx = lv.Context(local, stream)
        pctx.set_transfer_function(lv.scenes.standard_transfer_function(opacity=(0.1, 0.6)))
        pctx.set_option("ambient_occlusion_strength", 0.0)
        pctx.set_new_settings(extra_opts)
        if "avg_depth" in pw:
            pctx.set_option("b200_expected_avg_depth_complexity", pw["avg_depth"])
        if world > 1:
            pctx.set_tile_shard(rank, world, tile)
        pscene = pctx.create_scene(ppos, pattr, pseg, lv.scenes.LINE_WIDTH)
        pcam = lv.make_camera(pw["W"], pw["H"])
        pframe = torch.zeros((pw["H"], pw["W"], 4), dtype=torch.float32, device=dev)
        res, gat, pst = [], [], None
        for i in range(args.warmup + args.steps):
            pst = pctx.render_ppll(pscene, pcam, pw["max_frags"], "priority_queue", 0, out=pframe, stats=True)[1]
            if i >= args.warmup:
                res.append(pst["ms_resolve"]); gat.append(pst["ms_gather"])
        n_own_p = len(pctx.owned_tiles(pw["W"], pw["H"]))
        pc = torch.tensor([pst["frags_sorted"], float(np.mean(res)), float(np.mean(gat)), pst["frags_generated"], pst["frags_dropped"]],
                          dtype=torch.float64, device=dev)
        if world > 1:
            fs = pc[[0, 3, 4]].clone(); dist.all_reduce(fs)
            tm = pc[[1, 2]].clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            pc = torch.stack([fs[0], tm[0], tm[1], fs[1], fs[2]])
        frags, res_ms, gat_ms, gen, dropped = [float(x) for x in pc.tolist()]
        npx = pw["W"] * pw["H"]
        pbytes = 12 * pst["frags_sorted"] + 20 * (n_own_p * tile * tile if world > 1 else npx)
        out = {"workload": pw["desc"], "metric": "Mfrags/s sorted (PPLL resolve)", "value": frags / (res_ms * 1e-3) / 1e6,
               "unit": "Mfrags/s", "frags_sorted": frags, "frags_dropped": dropped, "ms_resolve": res_ms, "ms_gather": gat_ms,
               "gather_Mfrags_per_s": gen / (gat_ms * 1e-3) / 1e6, "max_depth_complexity": pst["max_depth_complexity"],
               "roofline": {"bound": "hbm", "achieved": pbytes / (float(np.mean(res)) * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": pbytes / (float(np.mean(res)) * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                            "kernel": "k_ppll_resolve", "bytes": "12 B/fragment + 20 B/pixel (SURVEY 8d), rank 0's share"}}
        if with_cpu and rank == 0 and not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline_ppll(pw, ppos, pattr, pseg, (960, 540))
        pscene.close(); pctx.close()
        del pframe
        torch.cuda.empty_cache()
        return out

    ppll_results = {}
    for i, name in enumerate(ppll_names):
        ppll_results["ppll" if i == 0 else "ppll_" + name] = measure_ppll(PPLL_WORKLOADS[name], with_cpu=(i == 0))

    # ---- the collective alone (N > 1): pack + all_gather + unpack on rank 0, CUDA events, max over ranks
    gather_ms = None
    if fg is not None:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        g0.record()
        for _ in range(max(3, args.steps)):
            if pf is not None:
                pf.fence()
            else:
                fg.gather(frame, assemble_on=(0,))
        g1.record(); torch.cuda.synchronize()
        gm = torch.tensor([g0.elapsed_time(g1) / max(3, args.steps)], dtype=torch.float64, device=dev)
        dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        gather_ms = float(gm.item())
    # per-rank time of the dominant kernel (load balance of the tile shards)
    krank = torch.tensor([k_ms], dtype=torch.float64, device=dev)
    if world > 1:
        kall = [torch.zeros_like(krank) for _ in range(world)]
        dist.all_gather(kall, krank)
        k_ms_ranks = [float(t.item()) for t in kall]
    else:
        k_ms_ranks = [k_ms]

    if rank == 0:
        value = tot_rays / (ms * 1e-3) / 1e6
        line = {
            "metric": "Mrays/s (tube+RTAO)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "frame": [W, H], "segments": int(info["n_seg"]), "bvh_nodes": int(info["n_nodes"]),
                       "scene_bytes": int(scene_bytes), "l2": "inputs larger than L2 (segments + BVH = %.2f GB)" % (scene_bytes / 1e9)
                       if scene_bytes > 200e6 else "scene fits L2; the 126 MB L2 is not flushed between frames",
                       "parallelism": ("tile-sharded x%d (64x64 tiles, Morton round-robin); " % world +
                                       ("frame assembled on rank 0 by NVLink peer stores from the frame kernels, 1-element all_reduce as fence"
                                        if peer else "1 NCCL all_gather/frame + unpack on rank 0")) if world > 1 else "single GPU",
                       "rays_per_step": tot_rays, "rays_primary": tot_rp, "rays_ao": tot_ra, "T_per_ray": tot_T / tot_rays, "I_per_ray": tot_I / tot_rays,
                       "scene_upload_and_bvh_build_s": upload_build_s, "bvh_build_ms": info["build_ms"], "bvh_build_ms_first_in_process": first_build_ms,
                       **({"options": extra_opts} if extra_opts else {})},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": ncu_traffic("k_rtao_rays_q", args.workload) if world == 1 else None, "algorithmic_bytes_per_launch": my_ao_bytes,
                         "kernel": "k_rtao_rays_q", "kernel_ms": k_ms, "peak_source": peak_src,
                         "bytes": "64 B x T + 32 B x I + 4 B per AO ray (SURVEY 8d); T/ray %.2f, I/ray %.2f over %d AO rays (rank 0)"
                                  % (st["ao_traversal_steps"] / max(st["rays_ao"], 1), st["ao_intersections"] / max(st["rays_ao"], 1), st["rays_ao"])},
            "e2e": {"value": tot_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": ctypes.sizeof(lv.LvCamera), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "note": "lv_render_tubes with a pinned HOST framebuffer: lv_camera struct in (the scene is resident, like the reference's cached render data), RGBA32F frame out"
                    if pf is None else "every rank: lv_render_tubes into rank 0's peer frame + fence; rank 0: whole RGBA32F frame D2H into pinned host memory; lv_camera struct in on every rank"},
            # k_rtao_primary, k_rtao_rays_q, k_rtao_reduce, k_tubes (+ tile pack / unpack kernels in all_gather mode)
            "gpu_launches": (4 + (2 + (world - 1) if (world > 1 and not peer) else 0)) * args.steps,
            "clocks": clocks,
        }
        line.update(ppll_results)
        if e2e8 is not None:
            line["e2e_rgba8"] = e2e8
        if gather_ms is not None:
            line["config"]["assemble_ms"] = gather_ms          # the frame fence (peer mode) or pack + all_gather + unpack alone, max over ranks
        line["config"]["k_rtao_rays_ms_per_rank"] = k_ms_ranks  # tile-shard load balance of the dominant kernel
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(wl, pos, attr, seg, tuple(args.ref_sample))
        emit(line)
    if pf is not None:
        torch.cuda.synchronize(); dist.barrier()
        pf.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
