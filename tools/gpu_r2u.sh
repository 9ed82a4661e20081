#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one_rank.py <<'PY'
import sys; sys.path.insert(0, ".")
import bench, torch, linevis_b200 as lv
wl = bench.WORKLOADS["config5"]
dev = torch.device("cuda", 0)
pos, attr, seg = bench.generate(wl["gen"], dev)
cam = lv.make_camera(wl["W"], wl["H"])
frame = torch.zeros((wl["H"], wl["W"], 4), dtype=torch.float32, device=dev)
ctx = lv.Context(0)
ctx.set_transfer_function(lv.scenes.standard_transfer_function())
ctx.set_new_settings({"depth_cue_strength": 0.0, "ambient_occlusion_strength": 1.0, "ambient_occlusion_gamma": 1.0, "ambient_occlusion_samples_per_frame": 64,
  "ambient_occlusion_iterations": 1, "ambient_occlusion_radius": 0.1, "ambient_occlusion_distance_based": True, "use_jittered_primary_rays": True,
  "num_samples_per_frame": 1, "num_accumulated_frames": 1, "use_deterministic_sampling": False})
sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
ctx.set_tile_shard(0, 8, 64)
for _ in range(3):
    ctx.render_tubes(sc, cam, 0, out=frame, stats=False)
torch.cuda.synchronize()
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" --csv --log-file gpurun_out/r2u_launches_rank0of8.csv python /tmp/one_rank.py > gpurun_out/r2u.log 2>&1; echo rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2u_launches_rank0of8.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[-12:]:
    print(r[ki][:70], r[vi], r[ui])
PY
