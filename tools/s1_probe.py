"""Stage-1 timing probe (GPU box): a batch of labelled scans of the street sequence through
sgtd_extract_instances_batch, device-resident inputs.  python tools/s1_probe.py [nscans] [reps]
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgtd_b200 import capi, synth, synth_seq  # noqa: E402

nscans = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
w = synth_seq.make_street_world(4541, synth.BASE_SEED + 1)
sel = np.linspace(0, 4540, nscans).astype(int)
pts, labs, off = [], [], [0]
for i in sel:
    p, l = synth_seq.render_at(w, w["poses"][i], 10_000 + int(i), device=dev)
    pts.append(p); labs.append(l.to(torch.int32)); off.append(off[-1] + p.shape[0])
pts, labs, off = torch.cat(pts).contiguous(), torch.cat(labs).contiguous(), np.array(off, np.int64)
mgr = capi.STDescManager(device=0)
mgr.set_option("s1_replay", int(os.environ.get("S1_REPLAY", "0")))
for it in range(reps):
    mgr.set_option("s1_trace", int(os.environ.get("S1_TRACE", "0")) if it == reps - 1 else 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nodes, noff, ninst = mgr.extract_instances_ptr(pts.data_ptr(), labs.data_ptr(), off)
    torch.cuda.synchronize()
    print(f"rep {it}: {1e3 * (time.perf_counter() - t0):.2f} ms for {nscans} scans, {int(off[-1])} points, {len(nodes)} nodes", flush=True)
