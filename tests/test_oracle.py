"""CPU tests of the oracle (test infrastructure) against INDEPENDENT restatements:
numpy SVD / Kabsch, scipy cKDTree, and a vectorised numpy re-derivation of the
reference's vote rule.  The reference ships no golden vectors (parity unpinned,
SURVEY.md 8c); these checks plus tests/golden/ freeze the oracle's behaviour.
"""
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from sgtd_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def small(oracle_lib):
    cfg = synth.make_config(0, 60, 3)
    xyz, lab, off = cfg["db"]
    o = oracle_lib.Oracle()
    descs = []
    for f in range(off.shape[0] - 1):
        d = o.build(xyz[off[f]:off[f + 1]], lab[off[f]:off[f + 1]])
        o.add(d)
        descs.append(d)
    return cfg, o, descs


def test_jacobi_svd_matches_numpy(oracle_lib):
    rng = np.random.default_rng(0)
    for it in range(200):
        A = rng.normal(size=(3, 3)) * 10 ** rng.uniform(-3, 3)
        if it % 3 == 0:  # rank 2, like the covariance of two centred triangles
            A = A @ np.diag([1, 1, 0]) @ rng.normal(size=(3, 3))
        U, s, V = oracle_lib.jacobi_svd3(A)
        np.testing.assert_allclose(U @ np.diag(s) @ V.T, A, atol=1e-12 * max(1, np.abs(A).max()))
        np.testing.assert_allclose(s, np.linalg.svd(A, compute_uv=False), rtol=1e-10, atol=1e-12 * np.abs(A).max())
        assert s[0] >= s[1] >= s[2] >= 0
        np.testing.assert_allclose(U.T @ U, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(V.T @ V, np.eye(3), atol=1e-12)


def kabsch_numpy(src, ref):
    """R/src/STDesc.cpp:549-571 with numpy.linalg.svd."""
    sc, rc = src.mean(0), ref.mean(0)
    H = (src - sc).T @ (ref - rc)
    U, _, Vt = np.linalg.svd(H)
    V = Vt.T
    R = V @ U.T
    if np.linalg.det(R) < 0:
        R = V @ np.diag([1, 1, -1]) @ U.T
    return R, rc - R @ sc


def test_triangle_solver_matches_numpy_kabsch(oracle_lib):
    rng = np.random.default_rng(1)
    for _ in range(100):
        src = rng.uniform(-40, 40, (3, 3)).astype(np.float32)
        ang = rng.uniform(-np.pi, np.pi)
        Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
        ref = (src @ Rz.T + rng.uniform(-5, 5, 3) + rng.normal(0, 0.05, (3, 3))).astype(np.float32)
        a = np.zeros(1, oracle_lib.DESC_DTYPE)
        b = np.zeros(1, oracle_lib.DESC_DTYPE)
        a["vert"][0] = src.reshape(9)
        b["vert"][0] = ref.reshape(9)
        R, t = oracle_lib.triangle_solver(a, b)
        R2, t2 = kabsch_numpy(src.astype(np.float64), ref.astype(np.float64))
        np.testing.assert_allclose(R, R2, atol=1e-9)
        np.testing.assert_allclose(t, t2, atol=1e-8)
        assert abs(np.linalg.det(R) - 1) < 1e-12


def test_build_invariants_and_knn(small, oracle_lib):
    cfg, o, descs = small
    xyz, lab, off = cfg["db"]
    for f in (0, 7, 31):
        P = xyz[off[f]:off[f + 1]]
        L = lab[off[f]:off[f + 1]]
        d = descs[f]
        assert len(d) > 0
        # emission order is (i, m, n) lexicographic and unique
        seq = d["anchor"].astype(np.int64) * 10000 + d["m"].astype(np.int64) * 100 + d["n"]
        assert (np.diff(seq) > 0).all()
        assert ((d["m"] >= 1) & (d["m"] < d["n"]) & (d["n"] <= 9)).all()
        # sides sorted, within the gate, and equal to the vertex distances AB, AC, BC
        s = d["side"]
        assert (s[:, 0] <= s[:, 1]).all() and (s[:, 1] <= s[:, 2]).all()
        assert (s >= 0.5).all() and (s <= 50).all()
        A, B, C = d["vert"][:, 0:3].astype(np.float64), d["vert"][:, 3:6].astype(np.float64), d["vert"][:, 6:9].astype(np.float64)
        np.testing.assert_allclose(np.linalg.norm(A - B, axis=1), s[:, 0], rtol=1e-6)
        np.testing.assert_allclose(np.linalg.norm(A - C, axis=1), s[:, 1], rtol=1e-6)
        np.testing.assert_allclose(np.linalg.norm(B - C, axis=1), s[:, 2], rtol=1e-6)
        # dedup key unique: (int64)(float)(side*1000)
        key = (s * 1000).astype(np.float32).astype(np.int64)
        assert np.unique(key, axis=0).shape[0] == key.shape[0]
        # every vertex is a node of the scan and labels follow the vertex
        tree = cKDTree(P.astype(np.float64))
        for V, col in ((A, 0), (B, 1), (C, 2)):
            dist, idx = tree.query(V)
            assert (dist == 0).all()
            assert (L[idx] == d["lab"][:, col]).all()
        # anchor's vertices are among its 10 nearest neighbours (exact kNN, float64 cross-check)
        _, nn = tree.query(P.astype(np.float64), k=10)
        _, ia = tree.query(A); _, ib = tree.query(B); _, ic = tree.query(C)
        for j in range(0, len(d), 37):
            allowed = set(nn[d["anchor"][j]])
            assert {ia[j], ib[j], ic[j]} <= allowed


def votes_numpy(q, db, n_frames, rough=0.03):
    """Independent vectorised restatement of R/src/STDesc.cpp:351-420."""
    votes = np.zeros(n_frames, np.int64)
    dbk = np.floor(db["side"] + 0.5).astype(np.int64)  # (int)(x+0.5), x > 0
    dbcode = ((db["lab"][:, 0].astype(np.int64) & 15) << 8) | ((db["lab"][:, 1].astype(np.int64) & 15) << 4) | (db["lab"][:, 2] & 15)
    order = np.argsort(dbcode, kind="stable")
    codes, starts = np.unique(dbcode[order], return_index=True)
    groups = {int(c): order[a:b] for c, a, b in zip(codes, starts, list(starts[1:]) + [order.size])}
    dbside = db["side"]
    for d in q:
        s = d["side"]
        thr = np.sqrt((s * s).sum()) * rough
        code = ((int(d["lab"][0]) & 15) << 8) | ((int(d["lab"][1]) & 15) << 4) | (int(d["lab"][2]) & 15)
        grp = groups.get(code)
        if grp is None:
            continue
        cand = grp[np.abs(dbside[grp] - s).max(axis=1) < 3.0]
        if cand.size == 0:
            continue
        dist = np.sqrt(((db["side"][cand] - s) ** 2).sum(axis=1))
        hit = (dist < thr) & (db["frame"][cand] != d["frame"])
        for x in (-1, 0, 1):
            for y in (-1, 0, 1):
                for z in (-1, 0, 1):
                    cell = np.trunc(s + np.array([x, y, z])).astype(np.int64)  # toward zero -> duplicates kept
                    if not np.sqrt(((s - (cell + 0.5)) ** 2).sum()) < 1.5:
                        continue
                    m = hit & (dbk[cand] == cell).all(axis=1)
                    np.add.at(votes, db["frame"][cand][m], 1)
    return votes


def test_votes_against_numpy_restatement(small, oracle_lib):
    cfg, o, descs = small
    qx, ql, qo = cfg["queries"]
    db = np.concatenate(descs)
    F = o.current_frame_id
    for q in range(qo.shape[0] - 1):
        qd = o.build(qx[qo[q]:qo[q + 1]], ql[qo[q]:qo[q + 1]])
        assert (qd["frame"] == F).all()
        r = o.search(qd)
        v = votes_numpy(qd, db, F)
        assert (r["votes"] == v).all()
        assert r["stats"]["M"] == v.sum()
        # ranking rule: first index of the strict maximum, >= 5 votes, at most candidate_num
        order = sorted([f for f in range(F) if v[f] >= 5], key=lambda f: (-v[f], f))[:50]
        assert list(r["cands"]["frame"]) == order
        assert list(r["cands"]["votes"]) == [int(v[f]) for f in order]
        # match lists: grouped per candidate, ordered by (query descriptor, probe ordinal, DB index)
        for c in r["cands"]:
            sl = slice(c["match_off"], c["match_off"] + c["nmatch"])
            assert c["nmatch"] == c["votes"]
            assert (db["frame"][r["m_g"][sl]] == c["frame"]).all()
            key = list(zip(r["m_q"][sl], r["m_cell"][sl], r["m_g"][sl]))
            assert key == sorted(key)


def test_verify_semantics(small, oracle_lib):
    """candidate_verify (R/src/STDesc.cpp:462-547) re-derived with numpy on the oracle's match lists."""
    cfg, o, descs = small
    qx, ql, qo = cfg["queries"]
    db = np.concatenate(descs)
    qd = o.build(qx[qo[0]:qo[1]], ql[qo[0]:qo[1]])
    r = o.search(qd)
    assert r["n"] > 0
    best, best_score = -1, 0
    for c in r["cands"]:
        sl = slice(c["match_off"], c["match_off"] + c["nmatch"])
        a = qd[r["m_q"][sl]]["vert"].astype(np.float64).reshape(-1, 3, 3)
        b = db[r["m_g"][sl]]["vert"].astype(np.float64).reshape(-1, 3, 3)
        M = a.shape[0]
        skip = M // 50 + 1
        H = M // skip
        votes = []
        for h in range(H):
            R, t = kabsch_numpy(a[h * skip], b[h * skip])
            res = np.linalg.norm(a @ R.T + t - b, axis=2)
            votes.append(int((res < 3.0).all(axis=1).sum()))
        hbest = int(np.argmax(votes)) if votes else 0
        if votes and votes[hbest] >= 4:
            assert c["best_hyp"] == hbest
            R, t = kabsch_numpy(a[hbest * skip], b[hbest * skip])
            inl = np.nonzero((np.linalg.norm(a @ R.T + t - b, axis=2) < 3.0).all(axis=1))[0]
            assert c["score"] == len(inl) == c["ninlier"]
            assert (r["inl"][c["inlier_off"]:c["inlier_off"] + c["ninlier"]] == inl).all()
            np.testing.assert_allclose(c["R"].reshape(3, 3), R, atol=1e-9)
            np.testing.assert_allclose(c["t"], t, atol=1e-8)
        else:
            assert c["score"] == -1
        if c["score"] > best_score:
            best_score, best = c["score"], c["frame"]
    assert r["best"] == ((best, float(best_score)) if best_score > 0.4 else (-1, 0.0))


def test_empty_and_small_inputs(oracle_lib):
    o = oracle_lib.Oracle()
    with pytest.raises(ValueError):
        o.build(np.zeros((5, 3), np.float32), np.zeros(5, np.uint32))
    r = o.search(np.zeros(0, oracle_lib.DESC_DTYPE))
    assert r["n"] == -1 and r["best"] == (-1, 0.0)  # "No STDescs!"
    rng = np.random.default_rng(2)
    d = o.build(rng.uniform(-20, 20, (30, 3)).astype(np.float32), rng.integers(3, 12, 30).astype(np.uint32))
    r = o.search(d)  # empty DB
    assert r["n"] == 0 and r["best"] == (-1, 0.0)


def test_probe_duplicate_cells_are_counted_twice(oracle_lib):
    """(int)(s+inc) truncates toward zero: for s in [0.5,1) inc=-1 and inc=0 address the same
    cell and the reference counts every match in it twice (R/src/STDesc.cpp:359-361)."""
    o = oracle_lib.Oracle()
    d = np.zeros(1, oracle_lib.DESC_DTYPE)
    d["side"][0] = (0.3, 5.2, 5.3)  # below the 0.5 m gate on purpose: DB cell x = (int)(0.8) = 0
    d["lab"][0] = (5, 8, 10)
    d["frame"][0] = 0
    for _ in range(6):
        o.add(d)  # frames 0..5 hold the same descriptor (frame field stays 0 like the reference would)
    q = d.copy()
    q["frame"] = 6
    r = o.search(q)
    assert r["votes"][0] == 12  # 6 entries x 2 duplicate cells
    assert r["stats"]["M"] == 12


def test_golden_fixture_is_reproduced(oracle_lib):
    g = np.load(os.path.join(GOLDEN, "stage234_small.npz"))
    cfg = synth.make_config(int(g["config_index"]), int(g["n_keyframes"]), int(g["n_queries"]))
    xyz, lab, off = cfg["db"]
    qx, ql, qo = cfg["queries"]
    o = oracle_lib.Oracle()
    nd = []
    for f in range(off.shape[0] - 1):
        d = o.build(xyz[off[f]:off[f + 1]], lab[off[f]:off[f + 1]])
        o.add(d)
        nd.append(len(d))
    assert (np.array(nd) == g["db_desc_counts"]).all()
    for q in range(qo.shape[0] - 1):
        qd = o.build(qx[qo[q]:qo[q + 1]], ql[qo[q]:qo[q + 1]])
        assert qd.tobytes() == g[f"q{q}_descs"].tobytes()
        r = o.search(qd)
        assert (r["votes"] == g[f"q{q}_votes"]).all()
        for k in ("frame", "votes", "nmatch", "score", "best_hyp", "ninlier"):
            assert (r["cands"][k] == g[f"q{q}_cand_{k}"]).all(), k
        np.testing.assert_array_equal(r["cands"]["R"], g[f"q{q}_cand_R"])
        np.testing.assert_array_equal(r["cands"]["t"], g[f"q{q}_cand_t"])
        assert (r["m_g"] == g[f"q{q}_m_g"]).all() and (r["m_q"] == g[f"q{q}_m_q"]).all()
        assert (r["m_cell"] == g[f"q{q}_m_cell"]).all() and (r["inl"] == g[f"q{q}_inl"]).all()
