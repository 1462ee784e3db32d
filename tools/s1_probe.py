"""Stage-1 timing probe (debug helper): a batch of synthetic scans through sgtd_extract_instances_batch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgtd_b200 import capi, synth_scan
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
base = [synth_scan.make_scan(5000 + s) for s in range(8)]
scans = [base[s % 8] for s in range(ns)]
off = np.concatenate([[0], np.cumsum([p.shape[0] for p, _ in scans])]).astype(np.int64)
P = np.concatenate([p for p, _ in scans]); L = np.concatenate([l for _, l in scans])
m = capi.STDescManager(device=0)
for it in range(3):
    t0 = time.perf_counter(); nodes, noff, pi, ninst = m.extract_instances(P, L, off); dt = time.perf_counter() - t0
    print(f"gpu batch {ns} scans: {dt*1e3:.1f} ms -> {ns/dt:.1f} scans/s, nodes/scan {len(nodes)/ns:.1f}")
t0 = time.perf_counter(); nodes, noff, pi, ninst = m.extract_instances(*base[0]); print("gpu single scan %.1f ms" % ((time.perf_counter()-t0)*1e3))
