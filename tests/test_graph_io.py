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


def _T(pose12):
    return np.vstack([np.asarray(pose12, np.float64).reshape(3, 4), [0, 0, 0, 1]])


def _rand_pose(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return np.column_stack([q, rng.normal(0, 50, 3)]).reshape(12)


def test_pose_error_matches_compute_adj_rpe():
    """sgtd_pose_error against the formula of compute_adj_rpe (R/include/utility.hpp:110-123)."""
    rng = np.random.default_rng(8)
    for _ in range(50):
        gt, est = _rand_pose(rng), _rand_pose(rng)
        d = np.linalg.inv(_T(est)) @ _T(gt)
        t_ref = np.linalg.norm(d[:3, 3])
        r_ref = abs(np.degrees(np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))))
        t, r = capi.pose_error(gt, est)
        assert abs(t - t_ref) < 1e-9 and abs(r - r_ref) < 1e-6
    t, r = capi.pose_error(gt, gt)
    assert t < 1e-9 and r < 1e-5
    with pytest.raises(capi.SgtdError):
        capi.pose_error(gt, np.zeros(12))      # singular estimate


def test_localization_check_is_the_main_loops_success_test():
    """MAt = T_map[match] * [R|t]_loop * refinement  against  gt * BASE2OUSTER; success iff T < 5 m and
    R < 10 deg (R/src/semantic_graph_localization.cpp:724-750).  The expected values are formed here with
    numpy from the reference's expressions, not from the library."""
    rng = np.random.default_rng(9)
    n_ok = 0
    for i in range(40):
        mp, gt, ex, rf = _rand_pose(rng), _rand_pose(rng), _rand_pose(rng), _rand_pose(rng)
        if i % 2:
            rf = np.eye(4)[:3].reshape(12)           # GICP disabled: transformation = identity
        # a loop transform that reproduces gt * extrinsic up to a perturbation of growing size
        want = np.linalg.inv(_T(mp)) @ _T(gt) @ _T(ex) @ np.linalg.inv(_T(rf))
        ang = np.radians(0.5 * i)
        c, s_ = np.cos(ang), np.sin(ang)
        pert = np.array([[c, -s_, 0, 0.2 * i], [s_, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
        loop = want @ pert
        ok, t, r, est = capi.localization_check(mp, loop[:3, :3], loop[:3, 3], gt, refine12=rf, gt_extr12=ex)
        MAt = _T(mp) @ loop @ _T(rf)                 # transform_j1 * new_trans * transformation
        gt_test = _T(gt) @ _T(ex)                    # transform_test * BASE2OUSTER
        assert np.allclose(_T(est), MAt, atol=1e-9)
        d = np.linalg.inv(MAt) @ gt_test             # compute_adj_rpe(gt = transform_test, lo = MAt)
        t_ref = np.linalg.norm(d[:3, 3])
        r_ref = abs(np.degrees(np.arccos(np.clip((np.trace(d[:3, :3]) - 1) / 2, -1, 1))))
        assert abs(t - t_ref) < 1e-8 and abs(r - r_ref) < 1e-5      # acos is ill-conditioned at 0 deg
        assert ok == (t_ref < 5.0 and r_ref < 10.0)
        n_ok += ok
    assert 0 < n_ok < 40                        # both outcomes were exercised
    # an extrinsic on the ground-truth side is NOT the same as one on the estimate
    ok1, t1, _, _ = capi.localization_check(mp, np.eye(3), np.zeros(3), mp, gt_extr12=ex)
    ok2, t2, _, _ = capi.localization_check(mp, np.eye(3), np.zeros(3), mp, refine12=ex)
    assert t1 > 1.0 and abs(t1 - t2) < 1e-6     # same distance here, but from opposite sides: est*E vs gt*E
    ok, t, r, _ = capi.localization_check(mp, np.eye(3), np.zeros(3), mp)   # nothing extra, exact pose
    assert ok and t < 1e-9 and r < 1e-5


def test_recall_rank_is_the_main_loops_bookkeeping():
    """sort by fitness descending, first candidate within 10 m of the ground truth -> its position
    (R/src/semantic_graph_localization.cpp:603-646)."""
    rng = np.random.default_rng(10)
    n_map = 30
    mp = np.stack([_rand_pose(rng) for _ in range(n_map)])
    for it in range(30):
        nc = int(rng.integers(0, 12))
        c = np.zeros(nc, capi.CAND_DTYPE)
        c["frame"] = rng.choice(n_map, nc, replace=False)
        c["score"] = rng.integers(-1, 6, nc)
        gt = mp[int(rng.integers(n_map))].copy()
        gt[[3, 7, 11]] += rng.normal(0, 4, 3)
        rank, order = capi.recall_rank(c, mp, gt)
        want_order = sorted(range(nc), key=lambda i: (-int(c["score"][i]), i))
        assert list(order) == want_order
        want = -1
        for pos, i in enumerate(want_order):
            t, _ = capi.pose_error(gt, mp[c["frame"][i]])
            if t < 10.0:
                want = pos
                break
        assert rank == want
