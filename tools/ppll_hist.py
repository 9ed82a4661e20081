"""Distribution of per-pixel list lengths for a PPLL workload + per-kernel times (diagnostic)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
import linevis_b200 as lv
pw = bench.PPLL_WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config4"]
pos, attr, seg = bench.generate(pw["gen"], torch.device("cuda", 0))
ctx = lv.Context(0)
ctx.set_transfer_function(lv.scenes.standard_transfer_function(opacity=(0.1, 0.6)))
sc = ctx.create_scene(pos, attr, seg, lv.scenes.LINE_WIDTH)
cam = lv.make_camera(pw["W"], pw["H"])
frame = torch.zeros((pw["H"], pw["W"], 4), dtype=torch.float32, device="cuda")
for binned in (False, True):
    ctx.set_option("b200_ppll_binned_resolve", binned)
    for i in range(3):
        _, st = ctx.render_ppll(sc, cam, pw["max_frags"], "priority_queue", 0, out=frame)
    print("binned" if binned else "plain ", "gather %.2f ms resolve %.3f ms" % (st["ms_gather"], st["ms_resolve"]))
# list lengths from the head/next structure is expensive on the host; use the alpha-less count: walk on GPU via torch
got = ctx.ppll_read()
nxt = torch.from_numpy(got["nodes"]["next"].astype(np.int64)).cuda()
heads = torch.from_numpy(got["heads"].astype(np.int64)).cuda()
cur = heads.clone(); cnt = torch.zeros_like(heads)
NONE = 0xFFFFFFFF
while True:
    m = cur != NONE
    if not bool(m.any()):
        break
    cnt[m] += 1
    cur[m] = nxt[cur[m]]
c = cnt.cpu().numpy()
print("pixels", c.size, "non-empty", int((c > 0).sum()), "frags", int(c.sum()), "max", int(c.max()))
edges = [0, 1, 9, 17, 33, 65, 129, 257, 100000]
for a, b in zip(edges[:-1], edges[1:]):
    sel = (c >= a) & (c < b)
    print("len [%d,%d): pixels %8d (%.1f%%)  frags %10d (%.1f%%)" % (a, b, sel.sum(), 100.0 * sel.mean(), c[sel].sum(), 100.0 * c[sel].sum() / max(c.sum(), 1)))
