"""CPU tests of the host-side file formats (no GPU needed): graph JSON wire compatibility
with Graph::toJSON/fromJSON (R/include/Semantic_Graph.hpp:79-157) and KITTI scan files."""
import json
import os

import numpy as np
import pytest

from sgtd_b200 import capi


def test_graph_json_roundtrip_and_schema(tmp_path):
    rng = np.random.default_rng(5)
    nodes = capi.make_nodes(rng.normal(0, 30, (57, 3)).astype(np.float32), rng.integers(3, 12, 57))
    pose = rng.normal(0, 10, 12).astype(np.float32)
    path = str(tmp_path / "000123.json")
    capi.graph_write_json(path, nodes, pose)
    j = json.load(open(path))
    # the seven keys Graph::toJSON emits; the four the reference leaves empty stay empty
    assert sorted(j) == ["centers", "densitys", "edges", "nodes", "poses", "volumes", "weights"]
    assert j["edges"] == j["weights"] == j["volumes"] == j["densitys"] == []
    assert j["nodes"] == nodes["label"].tolist() and len(j["poses"]) == 12
    # float -> JSON double -> float is exact
    c = np.array(j["centers"], np.float64).astype(np.float32)
    assert (c == np.column_stack([nodes["x"], nodes["y"], nodes["z"]])).all()
    back, poses = capi.graph_read_json(path)
    assert back.tobytes() == nodes.tobytes() and (poses == pose).all()


def test_reads_reference_style_json(tmp_path):
    """A file as nlohmann::json would dump it: arbitrary key order / whitespace, extra keys,
    doubles in shortest form, exponents."""
    path = str(tmp_path / "ref.json")
    open(path, "w").write(
        '{ "weights": [], "nodes":[5,10, 11],\n "edges":[[0.0,1.0]], "path": "a\\"b", '
        '"centers":[[1.5,-2.25,0.10000000149011612],[3e1,4.0E-1,-5],[0,0,0]],'
        '"poses":[1,0,0,10.5,0,1,0,-3,0,0,1,0.25], "volumes":[], "densitys":[1.0,2.0,3.0] }')
    nodes, poses = capi.graph_read_json(path)
    assert nodes["label"].tolist() == [5, 10, 11]
    assert nodes["x"].tolist() == [1.5, 30.0, 0.0] and nodes["z"][0] == np.float32(0.1)
    assert poses[3] == 10.5 and poses[11] == 0.25
    with pytest.raises(capi.SgtdError) as e:
        capi.graph_read_json(str(tmp_path / "missing.json"))
    assert e.value.status == capi.E_IO          # readGraphFromFile throws here
    open(path, "w").write('{"nodes":[1,2],"centers":[[1,2,3]]}')
    with pytest.raises(capi.SgtdError):
        capi.graph_read_json(path)              # centers/nodes length mismatch


def test_kitti_scan_files(tmp_path):
    rng = np.random.default_rng(6)
    pts = rng.normal(0, 20, (1000, 4)).astype(np.float32)
    lab = (rng.integers(0, 20, 1000) | (rng.integers(0, 50, 1000) << 16)).astype(np.uint32)
    b, l = str(tmp_path / "000000.bin"), str(tmp_path / "000000.label")
    pts.tofile(b); lab.tofile(l)
    p2, l2 = capi.scan_read_kitti(b, l)
    assert p2.tobytes() == pts.tobytes() and (l2 == lab).all()
    lab[:-1].tofile(l)
    with pytest.raises(capi.SgtdError):
        capi.scan_read_kitti(b, l)              # assert(points.cols()==labels.size())
