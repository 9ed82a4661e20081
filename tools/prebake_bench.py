"""Throughput of the object-space AO prebaker ("RTAO (Prebaker)") on the 100 k-segment helix: Mrays/s of the baking iterations
(k_bake_setup + k_rtao_rays<BAKE> + k_rtao_reduce), CUDA-event timed inside the library, beside the oracle on the host cores
for a bounded sample.  Prints one JSON line."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linevis_b200 as lv
from linevis_b200 import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--iterations", type=int, default=8)
ap.add_argument("--spp", type=int, default=4)
ap.add_argument("--param-len", type=float, default=0.001)
ap.add_argument("--cpu-vertices", type=int, default=4000)
args = ap.parse_args()

d = scenes.helix_polylines()
ctx = lv.Context(0)
ctx.set_new_settings({"ambient_occlusion_mode": "RTAO (Prebaker)", "b200_prebaker_iterations": args.iterations + 2,
                      "b200_prebaker_samples_per_frame": args.spp, "b200_prebaker_param_segment_length": args.param_len})
sc = ctx.create_scene(d["pos"], d["attr"], d["seg"], scenes.LINE_WIDTH)
t0 = time.time()
sc.set_lines(d["pos"], d["tangent"], d["normal"], d["line_offsets"])
set_lines_s = time.time() - t0
t0 = time.time()
sc.ao_bake(2)                      # warm-up (also builds the parametrization on the host)
warm_s = time.time() - t0
st = sc.ao_bake(args.iterations)
info = sc.ao_read()
line = {"metric": "Mrays/s (AO prebaker)", "value": st["rays_ao"] / (st["ms_rtao"] * 1e-3) / 1e6, "unit": "Mrays/s",
        "iterations": args.iterations, "rays": st["rays_ao"], "ms_total": st["ms_rtao"], "ms_rays_kernel": st["ms_rtao_rays"],
        "T_per_ray": st["ao_traversal_steps"] / max(st["rays_ao"], 1), "I_per_ray": st["ao_intersections"] / max(st["rays_ao"], 1),
        "n_param_vertices": info["n_param"], "n_subdiv": info["n_subdiv"], "spp": args.spp, "segments": int(d["seg"].shape[0]),
        "set_lines_s": set_lines_s, "first_bake_call_s": warm_s, "factor_mean": float(info["factors"].mean())}
if args.cpu_vertices:
    from oracle import lvo
    try:
        o = lvo.Oracle("ref"); kind = "reference"
    except (FileNotFoundError, OSError):
        o = lvo.Oracle("own"); kind = "port"
    osc = o.scene(d["pos"], d["attr"], d["seg"], scenes.LINE_WIDTH)
    osc.set_lines(d["tangent"], d["normal"])
    sl = info["sampling_locations"][:args.cpu_vertices]
    t0 = time.time()
    f, ost = osc.ao_bake_iteration(sl, 0, radius=0.1, n_subdiv=info["n_subdiv"], spp=args.spp)
    dt = time.time() - t0
    line["cpu_baseline"] = {"value": ost["rays"] / dt / 1e6, "unit": "Mrays/s", "cores": o.num_threads(), "kind": kind,
                            "sample": "first %d parametrization vertices, 1 iteration, %d rays in %.2f s" % (len(sl), ost["rays"], dt)}
print(json.dumps(line))
