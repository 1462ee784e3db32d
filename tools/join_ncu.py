"""ncu target: one search per join formulation on the bench database (see tools/join_probe.py).
ncu --set full --clock-control none --import-source on -k regex:"k_vote" python tools/join_ncu.py [impl ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgtd_b200 import capi, synth  # noqa: E402

impls = [int(x) for x in sys.argv[1:]] or [0, 1]
nkf, nq = 100000, 1024
cfg = synth.make_config(3, nkf, nq)
xyz, lab, off = cfg["db"]
qx, ql, qo = cfg["queries"]
mgr = capi.STDescManager(device=0)
nodes = capi.make_nodes(xyz, lab)
for c0 in range(0, nkf, 8192):
    c1 = min(nkf, c0 + 8192)
    b = mgr.build(nodes, off[c0:c1 + 1], frame_ids=np.arange(c0, c1, dtype=np.uint32))
    mgr.add(b); b.free()
mgr.finalize()
qb = mgr.build(capi.make_nodes(qx, ql), qo)
for impl in impls:
    mgr.set_option("join_impl", impl)
    for _ in range(2):
        mgr.search(qb).free()
