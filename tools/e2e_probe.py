"""Wall-clock breakdown of one e2e step (debug helper, not a benchmark)."""
import ctypes, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgtd_b200 import capi, synth
nkf, nq = int(sys.argv[1]), int(sys.argv[2])
cfg = synth.make_config(3, nkf, nq)
xyz, lab, off = cfg["db"]; qx, ql, qo = cfg["queries"]
mgr = capi.STDescManager(device=0)
nodes = capi.make_nodes(xyz, lab)
for c0 in range(0, nkf, 8192):
    c1 = min(nkf, c0 + 8192)
    b = mgr.build(nodes, off[c0:c1 + 1], frame_ids=np.arange(c0, c1, dtype=np.uint32)); mgr.add(b); b.free()
mgr.finalize()
qn = capi.make_nodes(qx, ql)
q_dev = torch.from_numpy(qn.view(np.uint8).reshape(-1)).cuda()
q_pin = torch.from_numpy(qn.view(np.uint8).reshape(-1).copy()).pin_memory()
q_pin_np = q_pin.numpy().view(capi.NODE_DTYPE)
k = mgr.cfg.candidate_num
loops_pin = torch.empty(nq * 16, dtype=torch.uint8).pin_memory()
cands_pin = torch.empty(nq * k * 136, dtype=torch.uint8).pin_memory()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    for mode in ("dev", "host"):
        t0 = T(); qb = mgr.build(q_dev.data_ptr() if mode == "dev" else q_pin_np, qo)
        t1 = T(); res = mgr.search(qb)
        t2 = T()
        if mode == "host":
            capi.lib().sgtd_result_download(mgr._h, res.ptr, ctypes.c_void_p(loops_pin.data_ptr()), ctypes.c_void_p(cands_pin.data_ptr()))
        t3 = T(); st, tm = res.stats(); res.free(); qb.free(); t4 = T()
        print(it, mode, "build %.1f search %.1f (lib total %.1f) download %.1f free %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, tm["total_ms"], (t3-t2)*1e3, (t4-t3)*1e3))
