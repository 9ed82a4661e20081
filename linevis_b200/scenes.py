"""Synthetic line sets of BASELINE.json's configs and the host-side segment builder.

All sets are normalised like the reference's loaders (centred, longest AABB side 0.5;
src/Loaders/TrajectoryFile.cpp:106-124) and returned as a flat "segment soup":
    pos  float32 [n_pt, 3], attr float32 [n_pt], seg_idx uint32 [n_seg, 2]
"""
import math

import numpy as np

LINE_WIDTH = 0.002  # src/Loaders/DataSetList.hpp:46


def normalize_positions(pos):
    """normalizeTrajectoriesVertexPositions: translate to the AABB centre, scale longest side to 0.5."""
    mn, mx = pos.min(axis=0), pos.max(axis=0)
    scale = float(np.min(0.5 / np.maximum(mx - mn, 1e-30)))
    return ((pos - 0.5 * (mn + mx)) * scale).astype(np.float32)


def segments_from_polylines(pos, attr, line_offsets, min_tangent_length=1e-4):
    """Host mirror of LineDataFlow::getLinePassTubeAabbRenderData (src/LineData/LineDataFlow.cpp:2140-2236):
    central-difference tangents, points with |tangent| < 1e-4 are skipped, consecutive surviving points of a
    trajectory are joined by index pairs (i-1, i); trajectories with <= 1 surviving point vanish.
    Returns (pos, attr, seg_idx) with the skipped points removed."""
    pos = np.asarray(pos, np.float32)
    attr = np.asarray(attr, np.float32)
    line_offsets = np.asarray(line_offsets, np.int64)
    n = pos.shape[0]
    if n == 0:
        return pos.reshape(0, 3), attr.reshape(0), np.zeros((0, 2), np.uint32)
    line_id = np.repeat(np.arange(len(line_offsets) - 1), np.diff(line_offsets))
    first = np.zeros(n, bool)
    last = np.zeros(n, bool)
    lens = np.diff(line_offsets)
    nz = lens > 0
    first[line_offsets[:-1][nz]] = True
    last[line_offsets[1:][nz] - 1] = True
    nxt = np.minimum(np.arange(n) + 1, n - 1)
    prv = np.maximum(np.arange(n) - 1, 0)
    a = np.where(last[:, None], pos, pos[nxt])
    b = np.where(first[:, None], pos, pos[prv])
    tangent = a - b
    # float32 length exactly like glm::length
    tl = np.sqrt((tangent[:, 0] * tangent[:, 0] + tangent[:, 1] * tangent[:, 1]) + tangent[:, 2] * tangent[:, 2])
    single = (lens == 1)[line_id]
    keep = (tl >= np.float32(min_tangent_length)) & ~single
    # drop trajectories with exactly one surviving point
    kept_per_line = np.bincount(line_id[keep], minlength=len(lens))
    keep &= (kept_per_line > 1)[line_id]
    new_index = np.cumsum(keep) - 1
    kp = np.nonzero(keep)[0]
    same_line = line_id[kp[1:]] == line_id[kp[:-1]]
    seg = np.stack([new_index[kp[:-1]][same_line], new_index[kp[1:]][same_line]], axis=1).astype(np.uint32)
    return pos[keep], attr[keep], seg


def polyline_frames(pos, line_offsets):
    """Per-point tangent / normal of polylines exactly like LineDataFlow::getLinePassTubeAabbRenderData
    (src/LineData/LineDataFlow.cpp:2150-2181): central-difference tangent (one-sided at the ends), normalised; normal =
    Gram-Schmidt of the previous point's normal (first point: (1,0,0); fallbacks (0,1,0), (0,0,1) when nearly parallel)
    against the tangent.  The polylines must be free of degenerate points (see segments_from_polylines).  Vectorised
    over the lines, sequential along them.  Returns (tangent [n,3], normal [n,3]) float32."""
    pos = np.asarray(pos, np.float32)
    off = np.asarray(line_offsets, np.int64)
    n = pos.shape[0]
    tangent = np.zeros((n, 3), np.float32)
    normal = np.zeros((n, 3), np.float32)
    lens = np.diff(off)
    f32 = np.float32

    def dot(a, b):
        return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]

    def normalize(a):
        return a * (f32(1.0) / np.sqrt(dot(a, a)))[:, None]

    def cross(a, b):
        return np.stack([a[:, 1] * b[:, 2] - b[:, 1] * a[:, 2], a[:, 2] * b[:, 0] - b[:, 2] * a[:, 0],
                         a[:, 0] * b[:, 1] - b[:, 0] * a[:, 1]], axis=1)

    last = np.tile(np.array([1, 0, 0], f32), (len(lens), 1))
    for j in range(int(lens.max()) if len(lens) else 0):
        act = np.nonzero(lens > j)[0]
        i = off[act] + j
        nxt = np.where(j + 1 < lens[act], i + 1, i)
        prv = np.where(j > 0, i - 1, i)
        t = normalize(pos[nxt] - pos[prv])
        helper = last[act].copy()
        bad = np.sqrt(dot(cross(helper, t), cross(helper, t))) < f32(0.01)
        helper[bad] = np.array([0, 1, 0], f32)
        bad2 = bad & (np.sqrt(dot(cross(helper, t), cross(helper, t))) < f32(0.01))
        helper[bad2] = np.array([0, 0, 1], f32)
        nrm = normalize(helper - dot(helper, t)[:, None] * t)
        tangent[i], normal[i] = t, nrm
        last[act] = nrm
    return tangent, normal


def polylines_with_frames(pos, attr, line_offsets):
    """Segment soup + what the object-space AO prebaker needs: drops degenerate points like segments_from_polylines, then
    returns dict(pos, attr, seg, tangent, normal, line_offsets) for the surviving polylines."""
    pos = np.asarray(pos, np.float32)
    line_offsets = np.asarray(line_offsets, np.int64)
    p2, a2, seg = segments_from_polylines(pos, attr, line_offsets)
    # surviving polylines = maximal runs of consecutive segments (i, i + 1)
    if len(seg) == 0:
        raise ValueError("no segments")
    brk = np.nonzero(seg[1:, 0] != seg[:-1, 1])[0] + 1
    starts = np.concatenate([[0], brk])
    first_pt = seg[starts, 0].astype(np.int64)
    off = np.concatenate([first_pt, [p2.shape[0]]])
    assert first_pt[0] == 0 and np.all(np.diff(off) >= 2)
    tangent, normal = polyline_frames(p2, off)
    return dict(pos=p2, attr=a2, seg=seg, tangent=tangent, normal=normal, line_offsets=off.astype(np.uint64))


def helix_polylines(n_lines=400, n_points=251, seed=1001):
    """The helix set of helix_lines() as polylines with frames (polylines_with_frames)."""
    rng = np.random.default_rng(seed)
    u = rng.random(n_lines)
    k = np.arange(n_lines)[:, None]
    j = np.arange(n_points)[None, :]
    rho = 0.02 + 0.23 * k / max(n_lines - 1, 1)
    ang = 2.0 * math.pi * 6.0 / 250.0 * j + 2.0 * math.pi * u[:, None]
    y = -0.25 + (0.5 / 250.0) * j + 0.0 * k
    pos = np.stack([rho * np.cos(ang), y, rho * np.sin(ang)], axis=-1).reshape(-1, 3)
    attr = ((y + 0.25) / 0.5).reshape(-1)
    return polylines_with_frames(normalize_positions(pos), attr.astype(np.float32), np.arange(n_lines + 1) * n_points)


def helix_lines(n_lines=400, n_points=251, seed=1001):
    """Config 2: 100 k-segment synthetic helix (Tornado-like).  Line k: radius 0.02 + 0.23 k/(n-1), pitch 0.5/250 per
    step in y, angular step 2 pi 6/250, phase 2 pi u_k, u_k ~ U(0,1); attribute = normalised y."""
    rng = np.random.default_rng(seed)
    u = rng.random(n_lines)
    k = np.arange(n_lines)[:, None]
    j = np.arange(n_points)[None, :]
    rho = 0.02 + 0.23 * k / max(n_lines - 1, 1)
    ang = 2.0 * math.pi * 6.0 / 250.0 * j + 2.0 * math.pi * u[:, None]
    y = -0.25 + (0.5 / 250.0) * j + 0.0 * k
    pos = np.stack([rho * np.cos(ang), y, rho * np.sin(ang)], axis=-1).reshape(-1, 3)
    attr = ((y + 0.25) / 0.5).reshape(-1)
    offsets = np.arange(n_lines + 1) * n_points
    pos = normalize_positions(pos)
    return segments_from_polylines(pos, attr.astype(np.float32), offsets)


def random_segments(n_seg=1_000_000, seg_len=0.01, seed=2002):
    """Configs 3/4: p0 ~ U([-.25,.25]^3), p1 = p0 + 0.01 unit(N(0,I)); attribute ~ U(0,1) per point."""
    rng = np.random.default_rng(seed)
    p0 = rng.random((n_seg, 3)) * 0.5 - 0.25
    d = rng.standard_normal((n_seg, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p1 = p0 + seg_len * d
    pos = np.empty((2 * n_seg, 3), np.float32)
    pos[0::2], pos[1::2] = p0, p1
    attr = rng.random(2 * n_seg).astype(np.float32)
    seg = np.arange(2 * n_seg, dtype=np.uint32).reshape(n_seg, 2)
    return pos, attr, seg


def curl_noise_streamlines(n_lines=20_000, n_points=501, step=5e-4, seed=3003, modes_per_octave=12, device=None):
    """Config 5: Rayleigh-Benard-like streamlines.  Velocity = curl(psi), psi = base roll potential
    (0, 0, sin(4 pi x) sin(pi y)) + a 3-octave random vector potential in a 2:1:2 box, integrated with RK4.
    The noise potential is spectral (random Fourier modes per octave, analytic curl) instead of lattice gradient
    noise -- same statistics class, exactly divergence-free, reproducible from `seed`.  attribute = |v| normalised."""
    rng = np.random.default_rng(seed)
    ks, amps, phases = [], [], []
    for o in range(3):
        f = 2.0 ** o
        kk = rng.standard_normal((modes_per_octave, 3))
        kk = kk / np.linalg.norm(kk, axis=1, keepdims=True) * (2.0 * math.pi * f)
        ks.append(kk)
        amps.append(rng.standard_normal((modes_per_octave, 3)) * (0.35 / f ** 2))
        phases.append(rng.random(modes_per_octave) * 2.0 * math.pi)
    kvec = np.concatenate(ks)            # [M,3]
    avec = np.concatenate(amps)          # [M,3]
    phi = np.concatenate(phases)         # [M]
    kxa = np.cross(kvec, avec)           # curl of a sin(k.x+phi) = (k x a) cos(k.x+phi)

    if device is not None:
        # same integration on the GPU (torch, float64) -- used by bench.py so that 10 M segments take seconds, not a minute
        import torch
        tk, tkxa, tphi = (torch.from_numpy(a).to(device) for a in (kvec, kxa, phi))

        def tvel(p):
            x, y = p[:, 0], p[:, 1]
            base = torch.stack([math.pi * torch.sin(4 * math.pi * x) * torch.cos(math.pi * y),
                                -4 * math.pi * torch.cos(4 * math.pi * x) * torch.sin(math.pi * y),
                                torch.zeros_like(x)], dim=1) * 0.25
            return base + torch.cos(p @ tk.T + tphi[None, :]) @ tkxa

        p = torch.from_numpy(rng.random((n_lines, 3)) * np.array([2.0, 1.0, 2.0])).to(device)
        ttraj = torch.empty((n_points, n_lines, 3), dtype=torch.float64, device=device)
        tspeed = torch.empty((n_points, n_lines), dtype=torch.float64, device=device)
        ttraj[0] = p
        for i in range(1, n_points):
            k1 = tvel(p)
            k2 = tvel(p + 0.5 * step * k1)
            k3 = tvel(p + 0.5 * step * k2)
            k4 = tvel(p + step * k3)
            tspeed[i - 1] = torch.linalg.norm(k1, dim=1)
            p = p + (step / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
            ttraj[i] = p
        tspeed[-1] = torch.linalg.norm(tvel(p), dim=1)
        traj, speed = ttraj.cpu().numpy(), tspeed.cpu().numpy()
    else:
        def vel(p):
            x, y = p[:, 0], p[:, 1]
            base = np.stack([math.pi * np.sin(4 * math.pi * x) * np.cos(math.pi * y),
                             -4 * math.pi * np.cos(4 * math.pi * x) * np.sin(math.pi * y),
                             np.zeros_like(x)], axis=1) * 0.25
            c = np.cos(p @ kvec.T + phi[None, :])
            return base + c @ kxa

        p = rng.random((n_lines, 3)) * np.array([2.0, 1.0, 2.0])
        traj = np.empty((n_points, n_lines, 3), np.float64)
        speed = np.empty((n_points, n_lines), np.float64)
        traj[0] = p
        for i in range(1, n_points):
            k1 = vel(p)
            k2 = vel(p + 0.5 * step * k1)
            k3 = vel(p + 0.5 * step * k2)
            k4 = vel(p + step * k3)
            speed[i - 1] = np.linalg.norm(k1, axis=1)
            p = p + (step / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
            traj[i] = p
        speed[-1] = np.linalg.norm(vel(p), axis=1)
    pos = traj.transpose(1, 0, 2).reshape(-1, 3)
    attr = speed.T.reshape(-1)
    attr = (attr - attr.min()) / max(attr.max() - attr.min(), 1e-30)
    offsets = np.arange(n_lines + 1) * n_points
    pos = normalize_positions(pos)
    return segments_from_polylines(pos, attr.astype(np.float32), offsets)


def _srgb_to_linear(c):
    c = np.asarray(c, np.float64)
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def standard_transfer_function(K=256, opacity=(1.0, 1.0)):
    """Data/TransferFunctions/Standard.xml: 5 sRGB colour points interpolated in linear RGB, opacity ramp
    opacity[0] -> opacity[1].  Returns K linear RGBA float32 entries (the LUT sgl would upload; sgl itself is not
    in the reference tree, so the LUT resolution/filtering is fixed by include/linevis_b200.h)."""
    pts = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    cols = _srgb_to_linear(np.array([[59, 76, 192], [144, 178, 254], [220, 220, 220], [245, 156, 125], [180, 4, 38]]) / 255.0)
    x = np.linspace(0.0, 1.0, K)
    lut = np.empty((K, 4), np.float32)
    for c in range(3):
        lut[:, c] = np.interp(x, pts, cols[:, c])
    lut[:, 3] = opacity[0] + (opacity[1] - opacity[0]) * x
    return lut
