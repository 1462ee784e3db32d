"""Freeze the stage-1 oracle on one small synthetic scan -> tests/golden/stage1_small.npz
(includes the scan itself, so the fixture also pins the generator).
    python tests/golden/make_golden_stage1.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from sgtd_b200 import synth_scan  # noqa: E402

SEED, N_AZ = 4242, 450   # ~28k points

if __name__ == "__main__":
    pts, lab = synth_scan.make_scan(SEED, n_az=N_AZ)
    r = orc.extract_instances(pts, lab)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage1_small.npz"),
                        seed=SEED, n_az=N_AZ, points=pts, labels=lab, n_instances=r["n_instances"],
                        point_instance=r["point_instance"], node_label=r["node_label"], node_xyz=r["node_xyz"])
    print("points", pts.shape, "instances", r["n_instances"], "nodes", len(r["node_label"]))
