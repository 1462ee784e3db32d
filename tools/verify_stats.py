"""Statistics of candidate verification on a scaled-down bench world (CPU, oracle as data source):
how many (hypothesis, pair) tests pass a single-vertex pre-test, per-candidate inlier rates.
Used to size k_verify's staged evaluation.  Test infrastructure (imports oracle/)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import orc  # noqa: E402
from sgtd_b200 import synth  # noqa: E402

nkf = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = synth.make_config(3, nkf, nq)
xyz, lab, off = cfg["db"]
o = orc.Oracle()
o.build_add_many(xyz, lab, off)
# DB descriptors again, for the vertices
db = []
o2 = orc.Oracle()
for f in range(nkf):
    d = o2.build(xyz[off[f]:off[f + 1]], lab[off[f]:off[f + 1]])
    d["frame"] = f
    db.append(d)
db = np.concatenate(db)
qx, ql, qo = cfg["queries"]
tot = dict(pairs=0, evals=0, passA=0, passAB=0, inl=0, passC=0)
for q in range(nq):
    od = o.build(qx[qo[q]:qo[q + 1]], ql[qo[q]:qo[q + 1]])
    r = o.search(od)
    c = r["cands"]
    moff = np.concatenate([[0], np.cumsum(c["nmatch"])])
    for ci in range(r["n"]):
        M = int(c["nmatch"][ci])
        mq = r["m_q"][moff[ci]:moff[ci + 1]]
        mg = r["m_g"][moff[ci]:moff[ci + 1]]
        a = od["vert"][mq].astype(np.float64).reshape(M, 3, 3)
        b = db["vert"][mg].astype(np.float64).reshape(M, 3, 3)
        skip = M // 50 + 1
        H = M // skip
        pA = pAB = pI = pC = 0
        for h in range(H):
            R, t = orc.triangle_solver(od[mq[h * skip]], db[mg[h * skip]])
            res = np.einsum("ij,mvj->mvi", R, a) + t - b
            d2 = (res ** 2).sum(-1)
            okA = d2[:, 0] < 9.0
            okAB = okA & (d2[:, 1] < 9.0)
            ok = okAB & (d2[:, 2] < 9.0)
            cen = (np.einsum("ij,mj->mi", R, a.mean(1)) + t - b.mean(1))
            pC += int(((cen ** 2).sum(-1) < 9.0).sum())
            pA += int(okA.sum()); pAB += int(okAB.sum()); pI += int(ok.sum())
        tot["pairs"] += M; tot["evals"] += M * H; tot["passA"] += pA; tot["passAB"] += pAB; tot["inl"] += pI; tot["passC"] += pC
        if ci < 6 or ci == r["n"] - 1:
            print(f"q{q} c{ci} frame {c['frame'][ci]} M {M} H {H} score {c['score'][ci]} "
                  f"passA {pA / (M * H):.3f} passAB {pAB / (M * H):.3f} inl {pI / (M * H):.3f} cen {pC / (M * H):.3f}")
print({k: v for k, v in tot.items()})
print({k: round(v / tot["evals"], 4) for k, v in tot.items() if k != "pairs"})
