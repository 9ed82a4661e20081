""".binlines round trip (reference src/Loaders/BinLinesLoader.cpp:40-240) and datasets.json entries."""
import json
import struct

import numpy as np
import pytest

from linevis_b200 import binlines, scenes


def test_roundtrip_v2_and_layout(tmp_path):
    pos, attr, seg = scenes.helix_lines(7, 13)
    lines = binlines.polylines_from_segments(pos, attr, seg)
    assert len(lines) == 7 and all(len(p) == 13 for p, _ in lines)
    path = tmp_path / "helix.binlines"
    binlines.write_binlines(path, lines)
    raw = path.read_bytes()
    # header and first trajectory exactly as loadTrajectoriesFromBinLinesV1 reads them
    assert struct.unpack_from("<III", raw, 0) == (2, 7, 1)
    assert struct.unpack_from("<I", raw, 12)[0] == 13
    assert np.array_equal(np.frombuffer(raw, "<f4", 39, 16).reshape(13, 3), lines[0][0])
    assert len(raw) == 12 + 7 * (4 + 13 * 12 + 13 * 4) + 6 * 4
    got = binlines.read_binlines(path)
    assert got["version"] == 2 and got["vertices_normalized"] and got["attribute_names"] == []
    for (p, a), (q, b) in zip(lines, got["lines"]):
        assert np.array_equal(p, q) and np.array_equal(a[0], b[0])
    # and back to the segment soup the renderer consumes
    allp = np.concatenate([p for p, _ in got["lines"]]); alla = np.concatenate([a[0] for _, a in got["lines"]])
    offs = np.concatenate([[0], np.cumsum([len(p) for p, _ in got["lines"]])])
    p2, a2, s2 = scenes.segments_from_polylines(allp, alla, offs)
    assert np.array_equal(p2, pos) and np.array_equal(a2, attr) and np.array_equal(s2, seg)


def test_v1_and_bad_magic(tmp_path):
    lines = [(np.arange(12, dtype=np.float32).reshape(4, 3), [np.arange(4, dtype=np.float32), np.ones(4, np.float32)])]
    p = tmp_path / "a.binlines"
    binlines.write_binlines(p, lines, version=1)
    got = binlines.read_binlines(p)
    assert got["version"] == 1 and len(got["lines"][0][1]) == 2 and np.array_equal(got["lines"][0][0], lines[0][0])
    bad = tmp_path / "bad.binlines"
    bad.write_bytes(struct.pack("<I", 7))
    with pytest.raises(ValueError):
        binlines.read_binlines(bad)


def test_datasets_json_schema():
    txt = binlines.datasets_json([dict(name="B200 Helix 100k", filename="flow/config2.binlines", linewidth=0.002)])
    d = json.loads(txt)
    e = d["datasets"][0]
    assert e["type"] == "flow" and e["name"] == "B200 Helix 100k" and e["filenames"].endswith(".binlines") and e["linewidth"] == 0.002
