"""Freeze the oracle's outputs on a small seeded case -> tests/golden/stage234_small.npz.

The reference ships no golden vectors and cannot be compiled here (parity
unpinned); these fixtures pin the CPU restatement so that any later change to
the oracle or to the synthetic generator is caught, and give the GPU tests a
reference that does not need the oracle at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from sgtd_b200 import synth  # noqa: E402

CONFIG_INDEX, N_KEYFRAMES, N_QUERIES = 0, 80, 2


def main():
    cfg = synth.make_config(CONFIG_INDEX, N_KEYFRAMES, N_QUERIES)
    xyz, lab, off = cfg["db"]
    qx, ql, qo = cfg["queries"]
    o = orc.Oracle()
    out = dict(config_index=CONFIG_INDEX, n_keyframes=N_KEYFRAMES, n_queries=N_QUERIES)
    counts = []
    for f in range(N_KEYFRAMES):
        d = o.build(xyz[off[f]:off[f + 1]], lab[off[f]:off[f + 1]])
        o.add(d)
        counts.append(len(d))
    out["db_desc_counts"] = np.array(counts)
    for q in range(N_QUERIES):
        qd = o.build(qx[qo[q]:qo[q + 1]], ql[qo[q]:qo[q + 1]])
        r = o.search(qd)
        out[f"q{q}_descs"] = qd
        out[f"q{q}_votes"] = r["votes"]
        for k in ("frame", "votes", "nmatch", "score", "best_hyp", "ninlier", "R", "t"):
            out[f"q{q}_cand_{k}"] = r["cands"][k]
        for k in ("m_q", "m_cell", "m_g", "inl"):
            out[f"q{q}_{k}"] = r[k]
        out[f"q{q}_best"] = np.array(r["best"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage234_small.npz"), **out)
    print("wrote stage234_small.npz:", {k: getattr(v, "shape", v) for k, v in out.items() if "cand_frame" in k})


if __name__ == "__main__":
    main()
