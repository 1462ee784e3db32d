"""CPU tests of the stage-1 oracle (oracle/dcvc_oracle.cpp) against an independent,
per-point pure-Python restatement of clusterManager (R/include/cluster_manager.hpp:172-421)
and of gen_labels / gen_graphs (R/src/get_json.cpp:41-299), plus a frozen fixture."""
import math
import os

import numpy as np

from sgtd_b200 import synth_scan

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def dcvc_python(xyz, startR=0.35, deltaR=0.0004, deltaP=1.2, deltaA=1.2):
    """Literal per-point DCVC in Python (small clouds only).  Returns label_info."""
    n = len(xyz)
    polar = [(0.0, 0.0, 0.0)] * n
    minPitch = maxPitch = 0.0
    minPolar = maxPolar = 5.0
    for i, (x, y, z) in enumerate(np.asarray(xyz, np.float64)):
        r = math.sqrt((x * x + y * y) + z * z)
        pitch = math.asin(z / r) * 180.0 / math.pi
        a = math.atan2(y, x)
        az = a * 180 / math.pi if a > 0.0 else (a + 2 * math.pi) * 180 / math.pi
        if r >= 120.0 or r <= 0.5:
            continue
        minPitch, maxPitch = min(minPitch, pitch), max(maxPitch, pitch)
        minPolar, maxPolar = min(minPolar, r), max(maxPolar, r)
        polar[i] = (r, pitch, az)
    width = int(round(360.0 / deltaA) + 1)
    height = int((maxPitch - minPitch) / deltaP)
    bounds, rng, step = [], minPolar, 1
    while rng <= maxPolar:
        rng += startR - step * deltaR
        bounds.append(rng)
        step += 1
    P = len(bounds)

    def rnd(v):  # std::round: half away from zero
        return int(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))

    def idx(i):
        r, pitch, az = polar[i]
        pi = next((k for k in range(P) if r < bounds[k]), P - 1)
        return pi, rnd((pitch - minPitch) / deltaP), rnd(az / deltaA)

    def vox(a, p, t):
        return (a * (P + 1) + p) + t * (P + 1) * (width + 1)

    vmap = {}
    for i in range(n):
        p, t, a = idx(i)
        vmap.setdefault(vox(a, p, t), []).append(i)
    lab = [-1] * n
    count = 0
    for i in range(n):
        if lab[i] != -1:
            continue
        p, t, a = idx(i)
        neigh = []
        for z in range(t - 1, t + 2):
            if z < 0 or z > height:
                continue
            for y in range(p - 1, p + 2):
                if y < 0 or y > P:
                    continue
                for x in range(a - 1, a + 2):
                    ax = width - 1 if x < 0 else x
                    ax = 300 if ax > 300 else ax
                    neigh += vmap.get(vox(ax, y, z), [])
        for j in neigh:
            c, nb = lab[i], lab[j]
            if c != -1 and nb != -1 and c != nb:
                lab = [nb if s == c else s for s in lab]
            elif nb != -1:
                lab[i] = nb
            elif c != -1:
                lab[j] = c
        if lab[i] == -1:
            count += 1
            lab[i] = count
            for j in neigh:
                lab[j] = count
    return np.array(lab), (width, height, P)


def test_dcvc_matches_python_restatement(oracle_lib):
    rng = np.random.default_rng(21)
    for it in range(10):
        n = int(rng.integers(50, 700))
        k = int(rng.integers(1, 6))
        c = rng.uniform(-30, 30, (k, 3)) * np.array([1, 1, 0.15])
        xyz = (c[rng.integers(0, k, n)] + rng.normal(0, rng.uniform(0.1, 2.0), (n, 3)) * np.array([1, 1, 0.4])).astype(np.float32)
        if it == 3:
            xyz[:5] *= 100  # out-of-range points keep a zero polar record
        lab, cl, nc, grid = oracle_lib.dcvc(xyz, minSeg=5)
        plab, pgrid = dcvc_python(xyz)
        assert grid == pgrid
        assert (lab == plab).all()
        # clusters_ == the label groups of size >= minSeg (their ORDER is libstdc++'s, checked by the fixture)
        sizes = {l: int((lab == l).sum()) for l in np.unique(lab)}
        assert nc == sum(1 for s in sizes.values() if s >= 5)
        for l, s in sizes.items():
            ids = np.unique(cl[lab == l])
            assert len(ids) == 1 and (ids[0] >= 0) == (s >= 5)


def test_gen_labels_policies(oracle_lib):
    rng = np.random.default_rng(22)
    n = 3000
    xyz = rng.uniform(-20, 20, (n, 3)) * np.array([1, 1, 0.1])
    pts = np.column_stack([xyz, np.zeros(n)]).astype(np.float32)
    sem = rng.choice([0, 8, 9, 10, 12, 4, 19, 14], n)
    inst = np.zeros(n, np.int64)
    m = sem == 4
    inst[m] = rng.choice([0, 3, 7, 500], m.sum(), p=[0.4, 0.3, 0.29, 0.01])
    lab = (sem | (inst << 16)).astype(np.uint32)
    r = oracle_lib.extract_instances(pts, lab)
    pi = r["point_instance"]
    # skipped classes never get an instance; classes 9 and 10 are one instance each (all their points)
    for c in (0, 8, 19, 14):
        assert (pi[sem == c] == -1).all()
    for c in (9, 10):
        assert len(np.unique(pi[sem == c])) == 1 and pi[sem == c][0] >= 0
    # GT-instance branch: one instance per id with > 20 points, ascending id, id 0 included
    ids = [i for i in sorted(np.unique(inst[m])) if ((inst == i) & m).sum() > 20]
    got = [int(pi[(inst == i) & m][0]) for i in ids]
    assert got == sorted(got) and len(set(got)) == len(ids)
    assert (pi[m & ~np.isin(inst, ids)] == -1).all()
    # instance ids ascend with the class id: 4 < 9 < 10 < 12
    first = {c: pi[(sem == c) & (pi >= 0)].min() for c in (4, 9, 10, 12) if ((sem == c) & (pi >= 0)).any()}
    assert list(first.values()) == sorted(first.values())
    # nodes: label = node_map[class] in 3..12 only (class 9 has no map entry, 4 maps to 0), float32 sequential centroid
    assert set(r["node_label"]) <= {3, 5}
    for k, ii in enumerate(r["node_inst"]):
        sel = pts[pi == ii, :3]
        acc = np.zeros(3, np.float32)
        for p in sel:
            acc = (acc + p).astype(np.float32)
        assert (r["node_xyz"][k] == acc / np.float32(len(sel))).all()


def test_stage1_golden_fixture(oracle_lib):
    g = np.load(os.path.join(GOLDEN, "stage1_small.npz"))
    pts, lab = synth_scan.make_scan(int(g["seed"]), n_az=int(g["n_az"]))
    assert pts.tobytes() == g["points"].tobytes() and (lab == g["labels"]).all()
    r = oracle_lib.extract_instances(pts, lab)
    assert r["n_instances"] == int(g["n_instances"])
    assert (r["point_instance"] == g["point_instance"]).all()
    assert (r["node_label"] == g["node_label"]).all() and r["node_xyz"].tobytes() == g["node_xyz"].tobytes()
