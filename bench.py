#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native LineVis hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Tube workloads (config5 = default, config3, small): a "step" is one frame of the tube + RTAO path (lv_render_tubes: RTAO pass + tube
ray-gen) on BASELINE.json's headline configuration ("10 M curl-noise streamline segments, 3840x2160, tubes + 64-spp RTAO"), which fits
one GPU.  `value` = Mrays/s with everything resident in HBM; `e2e` = the same metric through the C ABI with HOST buffers (camera in,
framebuffer read back to pinned host memory inside the timed region).  The PPLL path is measured beside it ("ppll", "ppll_config4").

PPLL workloads (config4, config2): a "step" is one lv_render_ppll frame (clear + gather + resolve); `value` = Mfrags/s sorted over the
whole frame time (the resolve-only rate of SURVEY 8d is reported under config.resolve_only_Mfrags_per_s).

--impl reference times the reference's own CPU path for the same workload: the oracle drivers on the reference's madmann91/bvh library
(oracle/_ref, built from /root/reference) or, if that is absent, the oracle port -- all host cores, a bounded crop of the frame.
"""
import argparse
import csv
import ctypes
import io
import json
import math
import os
import shutil
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
TILE = 64
# ncu metrics of the live capture: DRAM traffic of the launch and the counters that name the kernel's real limiter
NCU_METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
               "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
               "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
               "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


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


def committed_traffic(kernel, workload):
    """Fallback for roofline.traffic when no live ncu capture is possible: the committed capture's figure (profiles/traffic.json),
    labelled as such by the caller."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    e = json.load(open(p)).get(kernel, {})
    return e.get("dram_bytes_per_launch") if e.get("workload", "config5") == workload else None


def parse_ncu_csv(text):
    """ncu --csv (details page) -> {short kernel name: metrics of the LAST captured launch}; byte metrics in bytes."""
    lines = text.splitlines()
    start = next((i for i, l in enumerate(lines) if l.startswith('"ID"')), None)
    if start is None:
        return {}
    out = {}
    for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        name = row.get("Kernel Name", "").split("(")[0].split("<")[0].replace("void ", "").replace("lv::", "").strip()
        try:
            val = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = (row.get("Metric Unit") or "").strip()
        if "bytes" in row["Metric Name"]:
            val *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
        out.setdefault(name, {}).setdefault(int(row["ID"]), {})[row["Metric Name"]] = val
    return {name: by_id[max(by_id)] for name, by_id in out.items()}      # the last launch: warm caches, like the timed frames


def live_ncu(kernel_regex, child_args, timeout_s=300):
    """One ncu pass over a child process that renders two frames of the workload with the shipped library (tools/profile_run.py):
    per kernel matching `kernel_regex`, the last captured launch's DRAM bytes and limiter counters.  Numbers measured under the
    profiler are never timings of record -- only byte counts and percentages are used.  Returns {} when ncu is unavailable / fails."""
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {}
    cmd = [ncu, "--metrics", ",".join(NCU_METRICS), "--clock-control", "none", "-k", "regex:" + kernel_regex, "--csv",
           sys.executable, os.path.join(ROOT, "tools", "profile_run.py")] + child_args
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env, cwd=ROOT)
    except (subprocess.TimeoutExpired, OSError):
        return {}
    res = {}
    for name, m in parse_ncu_csv(r.stdout).items():
        res[name] = {
            "dram_bytes": m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0),
            "limiter": {"l1_wavefront_pct": m.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                        "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "lanes_per_inst": m.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                        "dram_pct": m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        "occupancy_pct": m.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                        "warp_inst": m.get("smsp__inst_executed.sum"),
                        "l2_hit_pct": m.get("lts__t_sector_hit_rate.pct"), "l1_hit_pct": m.get("l1tex__t_sector_hit_rate.pct"),
                        "source": "ncu pass inside this bench run (child process, shipped library, same workload; not a timing)"}}
    return res


def crop_camera(wl_w, wl_h, sw, sh):
    """Centre crop of the wl_w x wl_h frame at the full frame's ray density: same eye, narrower field of view, sw x sh pixels."""
    import linevis_b200 as lv
    return lv.make_camera(sw, sh, fov_y=2.0 * math.atan(0.5 * sh / wl_h))


def load_oracle():
    from oracle import lvo
    try:
        o = lvo.Oracle("ref"); kind = "reference"
    except (FileNotFoundError, OSError):
        o = lvo.Oracle("own"); kind = "port"
    o.set_num_threads()          # torchrun exports OMP_NUM_THREADS=1 for nproc > 1: size the pool explicitly (all host cores)
    return o, kind


# ------------------------------------------------------------------------------------------------- CPU legs
def oracle_tubes_frame(sc, cam, ao_spp):
    """One frame of the tube + RTAO path on the oracle: (image, rays, seconds, stats)."""
    import linevis_b200 as lv
    from oracle import lvo
    opts = lvo.default_options(ao_strength=1.0, ao_spp=ao_spp, ao_jitter_primary=1, ao_use_distance=1)
    tf = lv.scenes.standard_transfer_function()
    t0 = time.time()
    ao, s1 = sc.render_rtao(cam, opts, 0)
    img, s2 = sc.render_tubes(cam, opts, tf, ao_tex=ao)
    dt = time.time() - t0
    rays = s1["rays_primary"] + s1["rays_ao"] + s2["rays"]
    return img, rays, dt, dict(T=(s1["T"] + s2["T"]) / max(rays, 1), I=(s1["I"] + s2["I"]) / max(rays, 1), rays_ao=s1["rays_ao"],
                               T_ao=s1["T_ao"] / max(s1["rays_ao"], 1), I_ao=s1["I_ao"] / max(s1["rays_ao"], 1), ao=ao)


def oracle_ppll_frame(o, sc, cam, max_frags, canonical_image=False):
    """One PPLL frame on the oracle (gather + frontToBackPQ resolve): (image, frags sorted, gather s, resolve s, gather dict)."""
    import linevis_b200 as lv
    from oracle import lvo
    tf = lv.scenes.standard_transfer_function(opacity=(0.1, 0.6))
    opts = lvo.default_options()
    t0 = time.time()
    g = sc.ppll_gather(cam, opts, tf, linked_list_size=max_frags * cam.width * cam.height)   # room for every list: nothing is dropped
    t1 = time.time()
    assert g["counter"] <= max_frags * cam.width * cam.height, "oracle fragment buffer overflow"
    img, st = lvo.ppll_resolve(o, cam, opts, g["heads"], g["nodes"], max_frags, 0, canonical=False)   # the reference's own frontToBackPQ: the timed one
    t2 = time.time()
    if canonical_image:   # the result-defining order (DESIGN.md rule 9: ties in depth are broken by colour), what the CUDA path is compared with
        img = lvo.ppll_resolve(o, cam, opts, g["heads"], g["nodes"], max_frags, 0, canonical=True)[0]
    return img, st["frags_sorted"], t1 - t0, t2 - t1, g


def bounded_crop(W, H, s_per_px, budget_s, max_crop):
    """Largest centre crop (same aspect, width a multiple of 16) whose CPU frame fits `budget_s`, given seconds per pixel."""
    px = max(budget_s / max(s_per_px, 1e-12), 64.0 * 36.0)
    sw = int(min(max_crop[0], W, math.sqrt(px * W / H))) // 16 * 16
    sw = max(sw, 64)
    return sw, sw * H // W


def run_reference(args, name):
    """The reference's own CPU path for this workload, on the host cores (rank 0 only; the other ranks exit at once)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import linevis_b200 as lv
    o, kind = load_oracle()
    ppll = name in PPLL_WORKLOADS
    wl = PPLL_WORKLOADS[name] if ppll else WORKLOADS[name]
    gen_device = None
    if wl["gen"][0] == "curl":   # 10 M RK4-integrated points: torch on the CPU with every host thread (torchrun exports OMP_NUM_THREADS=1), ~15 s instead of a minute
        import torch
        torch.set_num_threads(o.num_threads())
        gen_device = torch.device("cpu")
    pos, attr, seg = generate(wl["gen"], gen_device)
    t0 = time.time()
    sc = o.scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    build_s = time.time() - t0
    W, H = wl["W"], wl["H"]
    # bounded sample of the same workload: the central crop of the full-resolution frame at the full frame's ray density (same
    # eye, same spp); the rate is intensive and reported as measured, never as an extrapolated time.  The crop is sized from a
    # small probe frame so that the whole arm (warm-up + steps) stays near --ref-budget seconds whatever the core count is.
    pw_ = max((W // 16) // 16 * 16, 64)
    ph_ = pw_ * H // W
    probe_cam = crop_camera(W, H, pw_, ph_)
    t0 = time.time()
    if ppll:
        oracle_ppll_frame(o, sc, probe_cam, wl["max_frags"])
    else:
        oracle_tubes_frame(sc, probe_cam, wl["ao_spp"])
    per_px = (time.time() - t0) / (pw_ * ph_)
    n_frames = max(args.warmup + args.steps, 1)
    sw, sh = bounded_crop(W, H, per_px, args.ref_budget / n_frames, tuple(args.ppll_sample if ppll else args.ref_sample))
    cam = crop_camera(W, H, sw, sh)
    times, units, extra = [], 0, {}
    for i in range(args.warmup + args.steps):
        if ppll:
            _, units, tg, tr, g = oracle_ppll_frame(o, sc, cam, wl["max_frags"])
            dt = tg + tr
            extra = {"gather_s": tg, "resolve_s": tr, "frags_generated": int(g["counter"])}
        else:
            _, units, dt, st = oracle_tubes_frame(sc, cam, wl["ao_spp"])
            extra = {"T_per_ray": st["T"], "I_per_ray": st["I"], "T_per_ao_ray": st["T_ao"], "I_per_ao_ray": st["I_ao"]}
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = units / (ms * 1e-3) / 1e6
    metric, unit = ("Mfrags/s sorted (PPLL frame)", "Mfrags/s") if ppll else ("Mrays/s (tube+RTAO)", "Mrays/s")
    sample = "%dx%d centre crop of the %dx%d frame at the frame's ray density, %d %s per step; BVH %s built in %.1f s (not timed)" % (
        sw, sh, W, H, units, "fragments sorted" if ppll else "rays", o.lib.lvo_backend_name().decode(), build_s)
    emit({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": wl["desc"], "sample": sample, **extra},
        "cpu_baseline": {"value": value, "unit": unit, "cores": o.num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------------- output
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


class Dist:
    """Rank bookkeeping + the few reductions the bench needs (no-ops on one GPU)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="sum"):
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def gather(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.dev)
        if self.world == 1:
            return [float(value)]
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(x.item()) for x in out]

    def timed(self, fn, steps, fork=None, join=None):
        """EXACTLY `steps` calls bracketed by barrier + synchronize on both sides; CUDA events; ms per step, max over ranks.
        fork / join: when `fn` enqueues on other streams than the current one, fork() makes them wait for the start event and join()
        makes the current stream wait for them before the end event is recorded."""
        self.barrier()
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record()
        if fork is not None:
            fork()
        for _ in range(steps):
            fn()
        if join is not None:
            join()
        e1.record()
        self.barrier()
        return self.reduce([e0.elapsed_time(e1) / steps], "max")[0]

    def timed_wall(self, fn, steps, finish=None):
        """The end-to-end legs return data to the host, so they are timed by the host clock around synchronised calls."""
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()                 # pipelined legs: the last frame's copy has landed inside the timed region
        self.torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        return self.reduce([ms], "max")[0]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


class Rgba8Pipeline:
    """The end-to-end step with the frame delivered as RGBA8 UNORM (the reference's sceneTexture format) and the device->host copy of
    frame i overlapping the kernels of frame i + 1.
      one GPU: `render(out)` is the C-ABI call with a pinned HOST frame; the library double-buffers (b200_async_delivery) and
               lv_synchronize() waits for the last copy;
      N GPUs:  every rank's frame kernels store rgba8 pixels into one of TWO peer frames on rank 0 (alternating), the fence is a 1-element
               all_reduce, rank 0 copies the finished frame to pinned host memory on a side stream while the next frame renders into the
               other peer frame; rank 0 joins fence k only after copy k-1 has left its buffer."""

    def __init__(self, D, ctxs, W, H, render, peer, streams=None):
        """ctxs: one context, or two (frames in flight: frame i is rendered by context i & 1 on streams[i & 1]); render(ctx, out)."""
        import linevis_b200  # noqa: F401
        from linevis_b200.sharding import PeerFrame
        torch = D.torch
        self.D, self.ctxs, self.render, self.i = D, list(ctxs), render, 0
        self.streams = streams if (streams is not None and len(self.ctxs) == 2) else None
        for c in self.ctxs:
            c.set_new_settings({"b200_frame_format": "rgba8", "b200_async_delivery": True})
        self.host = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
        self.host_np = [h.numpy().view(np.uint32) for h in self.host]
        self.pf = [PeerFrame(self.ctxs[0], W, H, D.rank, D.world, D.dev) for _ in range(2)] if peer else None
        if self.pf is not None and D.rank == 0:
            self.copy_stream = torch.cuda.Stream(device=D.dev)
            self.rendered = [torch.cuda.Event() for _ in range(2)]
            self.copied = [None, None]

    def _frame(self, j):
        torch, ctx = self.D.torch, self.ctxs[j % len(self.ctxs)]
        if self.pf is None:
            self.render(ctx, self.host_np[j])            # returns once the copy is enqueued
            return
        self.render(ctx, self.pf[j].ptr)
        if self.D.rank == 0:
            main = torch.cuda.current_stream()
            if self.copied[j ^ 1] is not None:
                main.wait_event(self.copied[j ^ 1])      # the other frame is written again after this fence: its copy must be done
            self.pf[j].fence()
            self.rendered[j].record(main)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.rendered[j])
                self.host[j].copy_(self.pf[j].tensor_rgba8(), non_blocking=True)
                self.copied[j] = torch.cuda.Event(); self.copied[j].record(self.copy_stream)
        else:
            self.pf[j].fence()

    def step(self):
        j = self.i & 1
        self.i += 1
        if self.streams is None:
            self._frame(j)
        else:
            with self.D.torch.cuda.stream(self.streams[j]):
                self._frame(j)

    def finish(self):
        for c in self.ctxs:
            c.synchronize()
        if self.streams is not None:
            for s_ in self.streams:
                s_.synchronize()
        if self.pf is not None and self.D.rank == 0:
            self.copy_stream.synchronize()

    def close(self):
        self.finish()
        for c in self.ctxs:
            c.set_new_settings({"b200_frame_format": "rgba32f", "b200_async_delivery": False})
        if self.pf is not None:
            self.D.barrier()
            for p in self.pf:
                p.close()


def parallelism_note(world, peer, shard="tiles"):
    if world == 1:
        return "single GPU"
    if shard == "samples":
        return ("pixels tile-sharded x%d (%dx%d tiles, Morton round-robin), AO rays sharded by SAMPLE batch (every rank: spp / %d samples of every hit "
                "pixel); per frame: all-gather of the hit lists, all-to-all of the per-sample results, " % (world, TILE, TILE, world) +
                ("peer-store frame assembly + 1-element fence" if peer else "all_gather frame assembly"))
    return "tile-sharded x%d (%dx%d tiles, Morton round-robin); " % (world, TILE, TILE) + (
        "frame assembled on rank 0 by NVLink peer stores from the frame kernels, 1-element all_reduce as fence" if peer
        else "1 NCCL all_gather/frame + unpack on rank 0")


def opt_args(extra_opts):
    return (["--opt"] + ["%s=%s" % kv for kv in extra_opts.items()]) if extra_opts else []


# ------------------------------------------------------------------------------------------------- PPLL measurement
def measure_ppll(D, args, name, extra_opts, hbm_peak, headline):
    """One PPLL workload on this rank's tiles.  headline=False: the side measurement of the default (tube) line -- kernel times from the
    library's own CUDA events.  headline=True: additionally the whole-frame step timing, e2e with a pinned host frame, CPU baseline with
    a crop-parity check, live ncu traffic / limiter for both kernels."""
    import linevis_b200 as lv
    from linevis_b200.sharding import FrameGather, PeerFrame
    torch = D.torch
    pw = PPLL_WORKLOADS[name]
    W, H = pw["W"], pw["H"]
    pos, attr, seg = generate(pw["gen"])
    stream = torch.cuda.current_stream().cuda_stream

    def make_ctx():
        c = lv.Context(D.local, stream)
        c.set_transfer_function(lv.scenes.standard_transfer_function(opacity=(0.1, 0.6)))
        c.set_option("ambient_occlusion_strength", 0.0)
        if "avg_depth" in pw:
            c.set_option("b200_expected_avg_depth_complexity", pw["avg_depth"])
        c.set_new_settings(extra_opts)
        return c
    ctx = make_ctx()
    if D.world > 1:
        ctx.set_tile_shard(D.rank, D.world, TILE)
    scene = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
    cam = lv.make_camera(W, H)
    frame = torch.zeros((H, W, 4), dtype=torch.float32, device=D.dev)
    res, gat, clr, st = [], [], [], None
    for i in range(args.warmup + args.steps):
        st = ctx.render_ppll(scene, cam, pw["max_frags"], "priority_queue", 0, out=frame, stats=True)[1]
        if i >= args.warmup:
            res.append(st["ms_resolve"]); gat.append(st["ms_gather"]); clr.append(st["ms_clear"])
    n_own_px = len(ctx.owned_tiles(W, H)) * TILE * TILE if D.world > 1 else W * H
    frags, gen, dropped = D.reduce([st["frags_sorted"], st["frags_generated"], st["frags_dropped"]])
    res_ms, gat_ms, clr_ms = D.reduce([np.mean(res), np.mean(gat), np.mean(clr)], "max")
    # algorithmic bytes, this rank's share: gather 12 B node store + 8 B head/count RMW per fragment; resolve 12 B node read
    # per fragment + 4 B head + 16 B RGBA32F write per pixel
    gather_bytes = 20 * st["frags_generated"]
    resolve_bytes = 12 * st["frags_sorted"] + 20 * n_own_px

    def roof(kernel, nbytes, ms, note):
        a = nbytes / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak, "traffic": None, "kernel": kernel,
                "kernel_ms": ms, "algorithmic_bytes_per_launch": nbytes, "bytes": note}
    mode = ctx.get_option("b200_ppll_gather_mode")
    gather_kernel = "k_ppll_gather" if mode == "raycast" else "k_ppll_gather_raster"
    r_res = roof("k_ppll_resolve", resolve_bytes, float(np.mean(res)), "12 B/fragment + 20 B/pixel (SURVEY 8d), rank 0's share")
    r_gat = roof(gather_kernel, gather_bytes, float(np.mean(gat)), "20 B/fragment: 12 B node + 8 B head/count (SURVEY 8d), rank 0's share")
    out = {"workload": pw["desc"], "metric": "Mfrags/s sorted (PPLL resolve)", "value": frags / (res_ms * 1e-3) / 1e6, "unit": "Mfrags/s",
           "frags_sorted": frags, "frags_generated": gen, "frags_dropped": dropped, "ms_resolve": res_ms, "ms_gather": gat_ms, "ms_clear": clr_ms,
           "gather_Mfrags_per_s": gen / (gat_ms * 1e-3) / 1e6, "max_depth_complexity": st["max_depth_complexity"],
           "gather_mode": mode, "roofline": r_res, "roofline_gather": r_gat}
    if not headline:
        scene.close(); ctx.close()
        del frame
        torch.cuda.empty_cache()
        return out

    # ---- headline: whole frames
    peer = D.world > 1 and args.assemble == "peer"
    pf = PeerFrame(ctx, W, H, D.rank, D.world, D.dev) if peer else None
    fg = FrameGather(W, H, TILE, D.rank, D.world, D.dev, ctx=ctx) if (D.world > 1 and not peer) else None

    def step():
        if pf is not None:
            ctx.render_ppll(scene, cam, pw["max_frags"], "priority_queue", 0, out=pf.ptr, stats=False)
            pf.fence()
        else:
            ctx.render_ppll(scene, cam, pw["max_frags"], "priority_queue", 0, out=frame, stats=False)
            if fg is not None:
                fg.gather(frame, assemble_on=(0,))
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(D.local) if D.rank == 0 else None
    ms = D.timed(step, args.steps)
    clocks = sampler.stop() if sampler else None

    host_frame = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
    host_np = host_frame.numpy()

    def step_e2e():
        if pf is not None:
            ctx.render_ppll(scene, cam, pw["max_frags"], "priority_queue", 0, out=pf.ptr, stats=False)
            pf.fence()
            if D.rank == 0:
                host_frame.copy_(pf.tensor(), non_blocking=True)
            pf.fence()
            torch.cuda.current_stream().synchronize()
            return
        ctx.render_ppll(scene, cam, pw["max_frags"], "priority_queue", 0, out=host_np, stats=False)     # D2H inside, synchronises
        if fg is not None:
            fg.gather(frame, assemble_on=())
    for _ in range(2):
        step_e2e()
    e2e_ms = D.timed_wall(step_e2e, args.steps)
    # the headline e2e: RGBA8 delivery (packed in the resolve kernel's epilogue), read back while the next frame renders
    e2e8 = None
    if fg is None:
        pipe = Rgba8Pipeline(D, [ctx], W, H, lambda c_, out: c_.render_ppll(scene, cam, pw["max_frags"], "priority_queue", 0, out=out, stats=False), peer)
        for _ in range(3):
            pipe.step()
        pipe.finish()
        ms8 = D.timed_wall(pipe.step, args.steps, finish=pipe.finish)
        ok = bool(D.rank != 0 or (np.count_nonzero(pipe.host_np[0]) == W * H and np.array_equal(pipe.host_np[0], pipe.host_np[1])))
        e2e8 = {"value": frags / (ms8 * 1e-3) / 1e6, "unit": "Mfrags/s", "ms_per_step": ms8, "h2d_bytes_per_step": ctypes.sizeof(lv.LvCamera),
                "d2h_bytes_per_step": W * H * 4, "frames_complete_and_equal": ok,
                "note": "lv_render_ppll, frame delivered as RGBA8 UNORM (the reference's sceneTexture format, packed in the resolve epilogue) to pinned host memory; "
                        "the copy of frame i overlaps frame i + 1 (" + ("b200_async_delivery, lv_synchronize inside the timed region" if not peer else
                                                                      "two alternating peer frames on rank 0, side-stream D2H") + ")"}
        pipe.close()
    launches_per_step = 3 if mode != "raster_contiguous" else 6   # clear (memset) + gather + resolve (+ scan, fill)

    line = None
    if D.rank == 0:
        dominant, other = (r_gat, r_res) if r_gat["kernel_ms"] >= r_res["kernel_ms"] else (r_res, r_gat)
        line = {
            "metric": "Mfrags/s sorted (PPLL frame)", "value": frags / (ms * 1e-3) / 1e6, "unit": "Mfrags/s", "n_gpus": D.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": pw["desc"], "frame": [W, H], "segments": int(scene.info()["n_seg"]), "step": "lv_render_ppll: clear + gather + resolve (frontToBackPQ)",
                       "l2": "fragment buffer %.2f GB, larger than L2" % (12 * gen / 1e9) if 12 * gen > 200e6 else "fragment buffer fits L2; not flushed between frames",
                       "parallelism": parallelism_note(D.world, peer), "frags_sorted": frags, "frags_generated": gen, "frags_dropped": dropped,
                       "ms_clear": clr_ms, "ms_gather": gat_ms, "ms_resolve": res_ms, "resolve_only_Mfrags_per_s": frags / (res_ms * 1e-3) / 1e6,
                       "gather_mode": mode, "max_depth_complexity": st["max_depth_complexity"], **({"options": extra_opts} if extra_opts else {})},
            "roofline": dominant, "roofline_" + ("resolve" if dominant is r_gat else "gather"): other,
            "e2e_rgba32f": {"value": frags / (e2e_ms * 1e-3) / 1e6, "unit": "Mfrags/s", "h2d_bytes_per_step": ctypes.sizeof(lv.LvCamera), "d2h_bytes_per_step": W * H * 16,
                            "ms_per_step": e2e_ms, "note": "synchronous float delivery: lv_render_ppll with a pinned HOST framebuffer (RGBA32F out, lv_camera in; scene resident)" if pf is None else
                            "synchronous float delivery: every rank renders into rank 0's peer frame + fence; rank 0: whole RGBA32F frame D2H into pinned host memory"},
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        }
        line["e2e"] = e2e8 if e2e8 is not None else line["e2e_rgba32f"]
        if D.world == 1:
            if not args.no_ncu:
                cap = live_ncu("k_ppll_gather|k_ppll_resolve", ["--skip-tubes", "--ppll-workload", name] + opt_args(extra_opts))
                for r in (r_gat, r_res):
                    c = cap.get(r["kernel"])
                    if c:
                        r["traffic"] = c["dram_bytes"]; r["limiter"] = c["limiter"]
            if not args.no_cpu_baseline:
                o, kind = load_oracle()
                osc = o.scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
                sw, sh = args.ppll_sample
                sub = crop_camera(W, H, sw, sh)
                ref_img, ref_frags, tg, tr, g = oracle_ppll_frame(o, osc, sub, pw["max_frags"], canonical_image=True)
                pctx = make_ctx()
                psc = pctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
                mine, mst = pctx.render_ppll(psc, sub, pw["max_frags"], "priority_queue", int(g["counter"]) + 1024)
                nan = np.isnan(ref_img)
                delta = float(np.abs(mine[~nan] - ref_img[~nan]).max()) if np.array_equal(np.isnan(mine), nan) else float("inf")
                psc.close(); pctx.close()
                line["cpu_baseline"] = {"value": ref_frags / (tg + tr) / 1e6, "unit": "Mfrags/s", "cores": o.num_threads(), "kind": kind,
                                        "sample": "%dx%d centre crop of the frame at its ray density: %d fragments gathered in %.2f s + resolved (frontToBackPQ) in %.3f s"
                                                  % (sw, sh, g["counter"], tg, tr), "resolve_only_Mfrags_per_s": ref_frags / tr / 1e6}
                line["parity_max_abs_delta"] = delta
                line["parity"] = {"max_abs_delta": delta, "fragment_counter_equal": bool(mst["frags_generated"] == g["counter"]),
                                  "frags_sorted_equal": bool(mst["frags_sorted"] == ref_frags), "tolerance": 1e-3,
                                  "what": "CUDA vs oracle (%s) on the CPU leg's %dx%d crop: resolved RGBA32F frame, fragment counter, fragments sorted" % (kind, sw, sh)}
    if pf is not None:
        D.barrier()
        pf.close()
    scene.close(); ctx.close()
    return line


# ------------------------------------------------------------------------------------------------- tube + RTAO measurement
def measure_tubes(D, args, name, extra_opts, hbm_peak, peak_src, ppll_names):
    import linevis_b200 as lv
    from linevis_b200.sharding import FrameGather, PeerFrame
    torch = D.torch
    rank, world, dev = D.rank, D.world, D.dev
    wl = WORKLOADS[name]
    W, H = wl["W"], wl["H"]
    pos, attr, seg = generate(wl["gen"], dev)
    stream = torch.cuda.current_stream().cuda_stream
    settings = {
        "depth_cue_strength": 0.0, "ambient_occlusion_strength": 1.0, "ambient_occlusion_gamma": 1.0,
        "ambient_occlusion_samples_per_frame": wl["ao_spp"], "ambient_occlusion_iterations": 1, "ambient_occlusion_radius": 0.1,
        "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True,
        "num_samples_per_frame": 1, "num_accumulated_frames": 1, "use_deterministic_sampling": False}
    def make_tube_ctx(cuda_stream):
        c = lv.Context(D.local, cuda_stream)
        c.set_transfer_function(lv.scenes.standard_transfer_function())
        c.set_new_settings(settings)
        c.set_new_settings(extra_opts)
        if world > 1:
            c.set_tile_shard(rank, world, TILE)
        return c
    ctx = make_tube_ctx(stream)
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
    peer = world > 1 and args.assemble == "peer"
    fg = FrameGather(W, H, TILE, rank, world, dev, ctx=ctx) if world > 1 else None     # also the e2e leg's device-side collective
    pf = PeerFrame(ctx, W, H, rank, world, dev) if peer else None

    # --shard samples (N > 1): the AO rays are split by SAMPLE instead of by tile (lv_sao_*, sharding.SampleShards): every rank traces
    # spp / N samples of every hit pixel of the frame; two more collectives per frame (hit lists, per-sample results)
    ss = None
    if args.frames_in_flight == 0:
        args.frames_in_flight = 2 if world > 1 else 1
    if world > 1 and args.shard == "samples":
        from linevis_b200.sharding import SampleShards
        ss = SampleShards(ctx, rank, world, wl["ao_spp"], dev)

    def render_frame(out, stats=False):
        if ss is not None:
            return ss.render(scene, cam, 0, out, stats=stats)[1]
        return ctx.render_tubes(scene, cam, 0, out=out, stats=stats)[1]

    def step(stats=False):
        if pf is not None:
            # every rank's frame kernels store their tiles straight into rank 0's frame (NVLink peer stores); the fence is the frame's only collective
            st = render_frame(pf.ptr, stats)
            pf.fence()
            return st
        st = render_frame(frame, stats)
        if fg is not None:
            fg.gather(frame, assemble_on=(0,))   # the single collective of the frame: every rank's packed tile block -> all ranks; rank 0 assembles
        return st

    # warm-up (also yields the per-frame ray / T / I counts: the frame is deterministic)
    st = None
    for _ in range(args.warmup):
        st = step(True)
    tot_rays, tot_T, tot_I, tot_rp, tot_ra, ao_T, ao_I = D.reduce([
        st["rays_primary"] + st["rays_ao"], st["traversal_steps"], st["intersections"], st["rays_primary"], st["rays_ao"],
        st["ao_traversal_steps"], st["ao_intersections"]])

    # ---- timed region: exactly K steps, barrier + synchronize on both sides, CUDA events, max over ranks.
    # --frames-in-flight 2 (default on several GPUs): frames alternate between two contexts on two streams that share the scene, so frame i + 1's packet
    # kernels run in the tail of frame i's persistent AO stream (linevis_b200.sharding.FramesInFlight; every frame is still one complete
    # lv_render_tubes frame, the frames are independent: no temporal accumulation in this workload).  ms_one_frame_in_flight is the
    # same loop with a single context, i.e. the latency of a frame.
    fif = None
    ms_single = D.timed(step, args.steps) if (args.frames_in_flight == 2 and ss is None) else None
    if args.frames_in_flight == 2 and ss is None:
        from linevis_b200.sharding import FramesInFlight
        streams2 = [torch.cuda.Stream(device=dev) for _ in range(2)]
        ctxs2 = [make_tube_ctx(s_.cuda_stream) for s_ in streams2]
        frames2 = [torch.zeros((H, W, 4), dtype=torch.float32, device=dev) for _ in range(2)] if not peer else None
        pfs2 = [PeerFrame(ctxs2[j], W, H, rank, world, dev) for j in range(2)] if peer else None

        def render2(c, j):
            c.render_tubes(scene, cam, 0, out=(pfs2[j].ptr if pfs2 is not None else frames2[j]), stats=False)
            if pfs2 is not None:
                pfs2[j].fence()
            elif fg is not None:
                fg.gather(frames2[j], assemble_on=(0,))
        fif = FramesInFlight(ctxs2, streams2, render2)
        for _ in range(max(4, args.warmup)):
            fif.step()
        torch.cuda.synchronize()
    sampler = ClockSampler(D.local) if rank == 0 else None
    ms = D.timed(fif.step, args.steps, fork=fif.fork, join=fif.join) if fif is not None else D.timed(step, args.steps)
    clocks = sampler.stop() if sampler else None
    if fif is not None:
        # the pipelined frames are complete and identical to the single-context frame
        torch.cuda.synchronize()
        ref_frame = pf.tensor() if (pf is not None and rank == 0) else frame
        fif_equal = True
        if rank == 0 or not peer:
            for j in range(2):
                got = pfs2[j].tensor() if pfs2 is not None else frames2[j]
                fif_equal = fif_equal and bool(torch.equal(got.view(torch.int32), ref_frame.view(torch.int32)))

    # ---- dominant kernel (the AO ray stream) live timing for the roofline: CUDA events around that kernel inside the library
    kt = []
    for _ in range(max(3, args.steps)):
        kt.append(render_frame(frame, True)["ms_rtao_rays"])
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
    e2e_ms = D.timed_wall(step_e2e, args.steps)
    # ---- the headline e2e: the frame delivered as RGBA8 UNORM -- the reference's own sceneTexture format (TubeRayTracing.glsl:42,
    # src/Widgets/DataView.cpp:100-108) -- packed in k_tubes' epilogue, read back to pinned host memory while the next frame renders
    e2e8 = None
    if fg is None or peer:
        pipe = Rgba8Pipeline(D, ctxs2 if fif is not None else [ctx], W, H, lambda c_, out: c_.render_tubes(scene, cam, 0, out=out, stats=False), peer,
                             streams=streams2 if fif is not None else None)
        for _ in range(3):
            pipe.step()
        pipe.finish()
        ms8 = D.timed_wall(pipe.step, args.steps, finish=pipe.finish)
        ok = bool(rank != 0 or (np.count_nonzero(pipe.host_np[0]) == W * H and np.array_equal(pipe.host_np[0], pipe.host_np[1])))
        e2e8 = {"value": tot_rays / (ms8 * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms8, "h2d_bytes_per_step": ctypes.sizeof(lv.LvCamera),
                "d2h_bytes_per_step": W * H * 4, "frames_complete_and_equal": ok, "frames_in_flight": 2 if fif is not None else 1,
                "note": ("lv_render_tubes with a pinned HOST frame, b200_frame_format = rgba8 (packed in the tube kernel's epilogue), b200_async_delivery: "
                         "the library copies frame i on a second stream while frame i + 1 renders; lv_synchronize inside the timed region") if not peer else
                        ("every rank: lv_render_tubes (rgba8) into one of two alternating peer frames on rank 0 + fence; rank 0: D2H of the finished "
                         "frame into pinned host memory on a side stream while the next frame renders; last copy inside the timed region")}
        pipe.close()
    # bytes read back per step: the whole frame on rank 0 (single GPU, or peer assembly), else every rank's own tiles (rank 0's share is reported)
    d2h = (n_own * TILE * TILE if (world > 1 and pf is None) else W * H) * 16

    # ---- the collective alone (N > 1): the frame fence, or pack + all_gather + unpack on rank 0; CUDA events, max over ranks
    gather_ms = None
    if fg is not None:
        gather_ms = D.timed((lambda: pf.fence()) if pf is not None else (lambda: fg.gather(frame, assemble_on=(0,))), max(3, args.steps))
    k_ms_ranks = D.gather(k_ms)          # per-rank time of the dominant kernel (load balance of the tile shards)

    # ---- PPLL path (second half of the metric), rank-local, whole frame on one GPU unless sharded
    ppll_results = {}
    for i, n in enumerate(ppll_names):
        ppll_results["ppll" if i == 0 else "ppll_" + n] = measure_ppll(D, args, n, extra_opts, hbm_peak, headline=False)

    line = None
    if rank == 0:
        ao_kernel = "k_rtao_rays_q" if ctx.get_option("b200_ao_queue") == "true" else "k_rtao_rays"
        if ao_kernel == "k_rtao_rays_q" and all(ctx.get_option(k_) == "true" for k_ in ("b200_ao_packed", "b200_ao_wide", "b200_ao_raybuf")):
            ao_kernel = "k_rtao_rays_w"
        line = {
            "metric": "Mrays/s (tube+RTAO)", "value": tot_rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "frame": [W, H], "segments": int(info["n_seg"]), "bvh_nodes": int(info["n_nodes"]),
                       **({"frames_in_flight": 2, "ms_one_frame_in_flight": ms_single, "frames_in_flight_equal_to_single": fif_equal} if fif is not None else {}),
                       "scene_bytes": int(scene_bytes), "l2": "inputs larger than L2 (segments + BVH = %.2f GB)" % (scene_bytes / 1e9)
                       if scene_bytes > 200e6 else "scene fits L2; the 126 MB L2 is not flushed between frames",
                       "parallelism": parallelism_note(world, peer, args.shard) + ("; two frames in flight (two contexts on two streams share the scene)" if fif is not None else ""),
                       "rays_per_step": tot_rays, "rays_primary": tot_rp, "rays_ao": tot_ra,
                       "T_per_ao_ray": ao_T / max(tot_ra, 1), "I_per_ao_ray": ao_I / max(tot_ra, 1),
                       "primary_packet_steps_per_ray": (tot_T - ao_T) / max(tot_rp, 1), "primary_packet_records_per_ray": (tot_I - ao_I) / max(tot_rp, 1),
                       "counting": "AO rays: one step / one record per lane; camera rays are traced as 32-ray warp packets and counted once per packet",
                       "scene_upload_and_bvh_build_s": upload_build_s, "bvh_build_ms": info["build_ms"], "bvh_build_ms_first_in_process": first_build_ms,
                       **({"options": extra_opts} if extra_opts else {})},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": None, "algorithmic_bytes_per_launch": my_ao_bytes,
                         "kernel": ao_kernel, "kernel_ms": k_ms, "peak_source": peak_src,
                         "bytes": "64 B x T + 32 B x I + 4 B per AO ray (SURVEY 8d) with the kernel's OWN T and I: T/ray %.2f, I/ray %.2f over %d AO rays (rank 0); "
                                  "cache-agnostic bookkeeping, not a DRAM measurement -- see traffic / limiter"
                                  % (st["ao_traversal_steps"] / max(st["rays_ao"], 1), st["ao_intersections"] / max(st["rays_ao"], 1), st["rays_ao"])},
            "e2e_rgba32f": {"value": tot_rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": ctypes.sizeof(lv.LvCamera), "d2h_bytes_per_step": int(d2h),
                            "ms_per_step": e2e_ms, "note": "synchronous float delivery: lv_render_tubes with a pinned HOST framebuffer, lv_camera struct in (the scene is resident, like the reference's cached render data), RGBA32F frame out"
                            if pf is None else "synchronous float delivery: every rank renders into rank 0's peer frame + fence; rank 0: whole RGBA32F frame D2H into pinned host memory"},
            # k_rtao_primary, AO ray stream, k_rtao_reduce, k_tubes (+ tile pack / unpack kernels in all_gather mode)
            "gpu_launches": (4 + (2 + (world - 1) if (world > 1 and not peer) else 0)) * args.steps,
            "clocks": clocks,
        }
        line.update(ppll_results)
        line["e2e"] = e2e8 if e2e8 is not None else line["e2e_rgba32f"]
        if gather_ms is not None:
            line["config"]["assemble_ms"] = gather_ms          # the frame fence (peer mode) or pack + all_gather + unpack alone, max over ranks
        line["config"]["k_rtao_rays_ms_per_rank"] = k_ms_ranks  # tile-shard load balance of the dominant kernel
        if world == 1 and not args.no_ncu:
            # DRAM traffic + limiter counters of the dominant kernel from an ncu pass of this very run (child process, shipped library)
            cap = live_ncu("k_rtao_rays", ["--skip-ppll", "--workload", name] + opt_args(extra_opts))
            c = cap.get(ao_kernel)
            if c:
                line["roofline"]["traffic"] = c["dram_bytes"]; line["roofline"]["limiter"] = c["limiter"]
                line["roofline"]["traffic_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum, measured inside this run"
        if world == 1 and line["roofline"]["traffic"] is None:
            line["roofline"]["traffic"] = committed_traffic(ao_kernel, name)
            line["roofline"]["traffic_source"] = "STALE: committed capture profiles/traffic.json (no live ncu pass in this run)"
        if not args.no_cpu_baseline and world == 1:
            # CPU baseline on a centre crop + parity of the CUDA path against it on exactly that crop + T / I of the reference library's tree
            o, kind = load_oracle()
            t0 = time.time()
            osc = o.scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
            build_s = time.time() - t0
            sw, sh = args.ref_sample
            sub = crop_camera(W, H, sw, sh)
            ref_img, rays, dt, ost = oracle_tubes_frame(osc, sub, wl["ao_spp"])
            line["cpu_baseline"] = {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": o.num_threads(), "kind": kind,
                                    "sample": "%dx%d centre crop of the frame at its ray density, %d spp, %d rays in %.1f s; CPU BVH (%s) build %.1f s not timed; "
                                              "T/ray %.1f, I/ray %.1f" % (sw, sh, wl["ao_spp"], rays, dt, o.lib.lvo_backend_name().decode(), build_s, ost["T"], ost["I"])}
            mine_ao, ast = ctx.render_rtao(scene, sub, 0)
            mine, mst = ctx.render_tubes(scene, sub, 0)
            delta = float(np.abs(mine - ref_img).max())
            line["parity_max_abs_delta"] = delta
            line["parity"] = {"max_abs_delta": delta, "ao_image_bit_exact": bool(np.array_equal(mine_ao.view(np.uint32), ost["ao"].view(np.uint32))),
                              "ao_rays_equal": bool(ast["rays_ao"] == ost["rays_ao"]), "tolerance": 1e-3,
                              "what": "CUDA vs oracle (%s) on the CPU leg's %dx%d centre crop: RGBA32F frame and AO image" % (kind, sw, sh)}
            # SURVEY 8d's roofline wording: algorithmic bytes from the reference library's tree on the same ray set, so that a better tree of
            # ours RAISES the fraction.  Same crop, same AO rays: bytes per AO ray on the reference tree / on our tree scales the achieved rate.
            own_T, own_I = ast["ao_traversal_steps"] / max(ast["rays_ao"], 1), ast["ao_intersections"] / max(ast["rays_ao"], 1)
            scale = (64 * ost["T_ao"] + 32 * ost["I_ao"] + 4) / (64 * own_T + 32 * own_I + 4)
            line["roofline"]["frac_ref_tree"] = achieved * scale / hbm_peak
            line["roofline"]["ref_tree"] = {"T_per_ao_ray": ost["T_ao"], "I_per_ao_ray": ost["I_ao"], "own_T_per_ao_ray_same_crop": own_T,
                                            "own_I_per_ao_ray_same_crop": own_I, "bytes_ratio_ref_over_own": scale,
                                            "note": "T / I of the %s tree for the AO rays of the %dx%d crop (oracle Statistics) against this kernel's counters on the same rays; "
                                                    "frac_ref_tree = frac x that ratio: it rises when our tree needs fewer steps" % (o.lib.lvo_backend_name().decode(), sw, sh)}
    if fif is not None:
        D.barrier()
        if pfs2 is not None:
            for p_ in pfs2:
                p_.close()
    if pf is not None:
        D.barrier()
        pf.close()
    scene.close(); ctx.close()
    if fif is not None:
        for c_ in ctxs2:
            c_.close()
    return line


def main():
    global TILE
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config5", choices=list(WORKLOADS) + list(PPLL_WORKLOADS),
                    help="tube + RTAO workloads: %s; PPLL workloads (the frame is then the headline step): %s" % (", ".join(WORKLOADS), ", ".join(PPLL_WORKLOADS)))
    ap.add_argument("--ppll-workload", default="config2,config4",
                    help="tube workloads only: comma-separated PPLL workloads measured beside the tube path (%s) or 'none'; the first is reported under "
                         "\"ppll\", further ones under \"ppll_<name>\"" % ", ".join(PPLL_WORKLOADS))
    ap.add_argument("--ref-sample", type=int, nargs=2, default=[1920, 1080],
                    help="centre crop (pixels) of the frame the CPU legs render: ~90 M rays on config 5, about 5-10 s per step on 16 host cores; "
                         "--impl reference shrinks it to fit --ref-budget")
    ap.add_argument("--ppll-sample", type=int, nargs=2, default=[480, 270], help="centre crop of the PPLL headline's CPU baseline / parity leg")
    ap.add_argument("--ref-budget", type=float, default=80.0, help="--impl reference: seconds of CPU rendering for warm-up + steps together")
    ap.add_argument("--frames-in-flight", type=int, default=0, choices=[0, 1, 2],
                    help="tube + RTAO headline: 2 = frames alternate between two contexts / streams sharing the scene, so one frame's fence, launch "
                         "gaps and stream tail overlap the next frame's kernels (config 5: 4.02 vs 4.45 ms on 8 GPUs; nothing on one GPU: 27.56 vs 27.55 ms, "
                         "and the pipelined e2e leg gets noisier, 28.2-28.9 vs 27.7 ms); 1 = one context; 0 (default) = 2 on several GPUs, 1 on one.  "
                         "With 2 the line carries ms_one_frame_in_flight as well")
    ap.add_argument("--shard", default="tiles", choices=["tiles", "samples"],
                    help="N > 1, tube + RTAO: tiles = image tiles only (default); samples = tiles for the pixels, AO rays split by sample batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="skip the live ncu pass (roofline.traffic then comes from the committed capture, labelled stale)")
    ap.add_argument("--assemble", default="peer", choices=["peer", "allgather"],
                    help="N > 1: how rank 0 gets the whole frame -- 'peer': every rank's kernels store their tiles straight into rank 0's "
                         "frame over NVLink (lv_frame_alloc / lv_ipc_*), one 1-element all_reduce as frame fence; 'allgather': pack + NCCL "
                         "all_gather + unpack")
    ap.add_argument("--tile", type=int, default=0, help="N > 1: tile size of the image shards in pixels (multiple of 16; default %d)" % TILE)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="extra lv_set_option settings for A/B runs (e.g. b200_ao_qnodes=true, b200_ppll_reg_sort=true); recorded in config.options")
    args = ap.parse_args()
    extra_opts = dict(o.split("=", 1) for o in args.opt)
    if args.tile:
        TILE = args.tile
    if args.impl == "reference":
        run_reference(args, args.workload)
        return
    args.warmup = max(args.warmup, 3)
    ppll_names = [n for n in args.ppll_workload.split(",") if n and n != "none"]
    for n in ppll_names:
        if n not in PPLL_WORKLOADS:
            ap.error("unknown PPLL workload %r" % n)

    D = Dist()
    hbm_peak, peak_src = peaks()
    if args.workload in PPLL_WORKLOADS:
        line = measure_ppll(D, args, args.workload, extra_opts, hbm_peak, headline=True)
        if line is not None:
            line["roofline"]["peak_source"] = peak_src
    else:
        line = measure_tubes(D, args, args.workload, extra_opts, hbm_peak, peak_src, ppll_names)
    if line is not None:
        emit(line)
    D.close()


if __name__ == "__main__":
    main()
