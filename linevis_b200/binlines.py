""".binlines reader / writer and datasets.json entries.

With these the UNCHANGED LineVis loaders (reference src/Loaders/BinLinesLoader.cpp:40-151 read, :153-240 write) can open
exactly the synthetic benchmark inputs of this repository on a Vulkan machine and produce the reference frame for the
(here unverifiable) Vulkan parity check.

Format (little endian):
  u32 version (1 | 2); u32 numTrajectories; u32 numAttributes;
  per trajectory: u32 numPoints; float32[numPoints][3] positions; numAttributes x float32[numPoints]
  version 2 appends: u32 verticesNormalized; u32 hasAttributeNames [names]; u32 hasRibbonData [float32[n][3] per trajectory];
  u32 numOutlineIndices, numOutlineVertices, numOutlineNormals [+ the three arrays].
Attribute names go through sgl::BinaryWriteStream::write(std::string); sgl is not part of the reference tree, so files
are WRITTEN with hasAttributeNames = 0 (unambiguous) and names are only READ under the assumption "u32 length + bytes".
"""
import json
import struct

import numpy as np


def write_binlines(path, lines, vertices_normalized=True, version=2):
    """lines: iterable of (positions float32 [n, 3], attributes: list of float32 [n])."""
    lines = list(lines)
    n_attr = len(lines[0][1]) if lines else 0
    with open(path, "wb") as f:
        f.write(struct.pack("<III", version, len(lines), n_attr))
        for pos, attrs in lines:
            pos = np.ascontiguousarray(pos, "<f4").reshape(-1, 3)
            assert len(attrs) == n_attr
            f.write(struct.pack("<I", pos.shape[0]))
            f.write(pos.tobytes())
            for a in attrs:
                a = np.ascontiguousarray(a, "<f4").reshape(-1)
                assert a.shape[0] == pos.shape[0]
                f.write(a.tobytes())
        if version == 2:
            f.write(struct.pack("<III", 1 if vertices_normalized else 0, 0, 0))   # normalized, no names, no ribbons
            f.write(struct.pack("<III", 0, 0, 0))                                 # no simulation mesh outline


def read_binlines(path):
    """Returns dict(version, lines=[(positions [n,3], [attribute arrays])], vertices_normalized, attribute_names)."""
    buf = open(path, "rb").read()
    off = 0

    def u32():
        nonlocal off
        v = struct.unpack_from("<I", buf, off)[0]
        off += 4
        return v

    def f32(n):
        nonlocal off
        a = np.frombuffer(buf, "<f4", n, off).copy()
        off += 4 * n
        return a

    version = u32()
    if version not in (1, 2):
        raise ValueError("invalid magic number in %s" % path)   # BinLinesLoader.cpp:137-141
    n_traj, n_attr = u32(), u32()
    lines = []
    for _ in range(n_traj):
        n = u32()
        pos = f32(3 * n).reshape(n, 3)
        attrs = [f32(n) for _ in range(n_attr)]
        lines.append((pos, attrs))
    out = dict(version=version, lines=lines, vertices_normalized=False, attribute_names=[])
    if version == 2:
        out["vertices_normalized"] = u32() != 0
        if u32() != 0:
            for _ in range(n_attr):
                ln = u32()
                out["attribute_names"].append(buf[off:off + ln].decode("utf-8", "replace"))
                off += ln
        if u32() != 0:
            out["ribbon_directions"] = [f32(3 * len(p)).reshape(-1, 3) for p, _ in lines]
        ni, nv, nn = u32(), u32(), u32()
        out["outline_indices"] = np.frombuffer(buf, "<u4", ni, off).copy(); off += 4 * ni
        out["outline_vertices"] = f32(3 * nv).reshape(nv, 3)
        out["outline_normals"] = f32(3 * nn).reshape(nn, 3)
    return out


def polylines_from_segments(pos, attr, seg_idx):
    """Inverse of scenes.segments_from_polylines for export: consecutive index pairs (i, i+1) are chained into trajectories."""
    seg_idx = np.asarray(seg_idx)
    lines = []
    if len(seg_idx) == 0:
        return lines
    breaks = np.nonzero(seg_idx[1:, 0] != seg_idx[:-1, 1])[0] + 1
    for a, b in zip(np.concatenate([[0], breaks]), np.concatenate([breaks, [len(seg_idx)]])):
        ids = np.concatenate([seg_idx[a:b, 0], seg_idx[b - 1:b, 1]])
        lines.append((pos[ids], [attr[ids]]))
    return lines


def datasets_json(entries):
    """entries: list of dict(name, filename, linewidth, attributes) -> text of Data/LineDataSets/datasets.json (README.md:117-135)."""
    return json.dumps({"datasets": [{"type": "flow", "name": e["name"], "filenames": e["filename"],
                                     "linewidth": e.get("linewidth", 0.002), "attributes": e.get("attributes", "Attribute")}
                                    for e in entries]}, indent=4)
