"""Surface-area cost (sum of the areas of the wide nodes / root area) of the greedy 4-wide collapse against the OPTIMAL collapse of the same
child-pair tree (dynamic programme over the number of slots a subtree may occupy in its nearest wide ancestor), on the emulated library:
    python tools/collapse_cost.py curl|random [lbvh|sah|ploc]
Result (150 k segments): the optimum is 2.3 - 2.5 % below the greedy collapse for every builder -- not worth a second collapse kernel."""
import sys, numpy as np, importlib.util, time
sys.path.insert(0, "/root/repo")
import linevis_b200 as lv
from linevis_b200 import scenes
spec = importlib.util.spec_from_file_location("build_emu", "/root/repo/tests/emu/build_emu.py"); mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
c = lv.Context(0, lib_path=mod.build())
c.set_option("b200_ao_wide", False)
which = sys.argv[1] if len(sys.argv) > 1 else "curl"
if which == "curl":
    pos, attr, seg = scenes.curl_noise_streamlines(n_lines=300, n_points=501)
else:
    pos, attr, seg = scenes.random_segments(150000, 0.01, seed=5)
print("segments", seg.shape[0])
if len(sys.argv) > 2: c.set_option("b200_bvh_builder", sys.argv[2])
sc = c.create_scene(pos, attr, seg, scenes.LINE_WIDTH)
N = sc.bvh_nodes()
n = N.shape[0]
lref, rref = N["lref"].astype(np.int64), N["rref"].astype(np.int64)
def area(mn, mx):
    d = (mx - mn).astype(np.float64); return d[:,0]*d[:,1] + d[:,1]*d[:,2] + d[:,2]*d[:,0]
Al, Ar = area(N["lmin"], N["lmax"]), area(N["rmin"], N["rmax"])
leafL, leafR = (lref >> 31) & 1, (rref >> 31) & 1
# node area: from parent's child box. root area = union
A = np.zeros(n); 
rootmn = np.minimum(N["lmin"][0], N["rmin"][0]); rootmx = np.maximum(N["lmax"][0], N["rmax"][0])
A[0] = area(rootmn[None], rootmx[None])[0]
inner_l = np.where(leafL == 0)[0]; A[lref[inner_l]] = Al[inner_l]
inner_r = np.where(leafR == 0)[0]; A[rref[inner_r]] = Ar[inner_r]
# BFS order
order = []; level = [0]
while level:
    order.extend(level); nxt = []
    for x in level:
        if not leafL[x]: nxt.append(lref[x])
        if not leafR[x]: nxt.append(rref[x])
    level = nxt
order = np.array(order)
# greedy collapse cost: sum of area of wide roots
def greedy():
    tot = 0.0; nwide = 0; q = [0]
    while q:
        nq = []
        for x in q:
            tot += A[x]; nwide += 1
            ch = []  # (is_leaf, ref, area)
            ch.append((leafL[x], lref[x], Al[x])); ch.append((leafR[x], rref[x], Ar[x]))
            while len(ch) < 4:
                best = -1; ba = -1
                for i,(lf, r, a) in enumerate(ch):
                    if not lf and a > ba: ba = a; best = i
                if best < 0: break
                lf, r, a = ch.pop(best)
                ch.append((leafL[r], lref[r], Al[r])); ch.append((leafR[r], rref[r], Ar[r]))
            for lf, r, a in ch:
                if not lf: nq.append(r)
        q = nq
    return tot, nwide
INF = 1e300
f = np.zeros((n, 4))   # f[x][k], k=1..3
dec = np.zeros((n, 4), np.int8)
for x in order[::-1]:
    def F(isleaf, r, k): return 0.0 if isleaf else f[r][k]
    # g = best with 4 slots
    g = min(F(leafL[x], lref[x], j) + F(leafR[x], rref[x], 4 - j) for j in (1, 2, 3))
    f[x][1] = A[x] + g
    s2 = F(leafL[x], lref[x], 1) + F(leafR[x], rref[x], 1)
    f[x][2] = min(f[x][1], s2)
    s3 = min(F(leafL[x], lref[x], 1) + F(leafR[x], rref[x], 2), F(leafL[x], lref[x], 2) + F(leafR[x], rref[x], 1))
    f[x][3] = min(f[x][2], s3)
gt, gw = greedy()
print("greedy: sum area / root area = %.3f, wide nodes %d" % (gt / A[0], gw))
print("optimal: sum area / root area = %.3f" % (f[0][1] / A[0]))
print("gain %.2f %%" % (100 * (1 - f[0][1] / gt)))
