#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_prebaker.py tests/test_gpu_parity.py -m gpu -q -k "prebak or stack_layouts" > gpurun_out/e1_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/e1_tests.log
timeout 600 python tools/sweep.py --leaf 1 --refill 24 --vote 12 --minb 8 10 12 --stack 0 1 2 > gpurun_out/e1_sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/e1_sweep.log
timeout 200 python - <<'PY' > gpurun_out/e1_build.log 2>&1
import sys, time; sys.path.insert(0, '.')
import torch, numpy as np, bench, linevis_b200 as lv
dev = torch.device("cuda", 0)
pos, attr, seg = bench.generate(bench.WORKLOADS["config5"]["gen"], dev)
d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (pos, attr, seg.view(np.int32))]
ctx = lv.Context(0)
for i in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    sc = ctx.create_scene(d[0], d[1], d[2], lv.scenes.LINE_WIDTH)
    torch.cuda.synchronize()
    print("create %d: wall %.1f ms, build_ms %.2f" % (i, 1e3 * (time.time() - t0), sc.info()["build_ms"]), flush=True)
    sc.close()
PY
cat gpurun_out/e1_build.log
