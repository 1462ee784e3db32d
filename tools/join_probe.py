"""Experiment driver (GPU box): builds the bench database once and times the search stages under
different handle options.  python tools/join_probe.py [keyframes] [queries]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgtd_b200 import capi, synth  # noqa: E402


def main():
    nkf = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    combos = sys.argv[3:] or ["", "join_hint=1", "join_parts=2", "join_parts=2,join_hint=1", "join_parts=4",
                              "join_parts=4,join_hint=1", "join_parts=2,join_groups=3", "join_parts=2,join_groups=6",
                              "join_hint=1,join_groups=6", "join_hint=1,join_groups=5", "verify_impl=1"]
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
    names = ("join_impl", "join_groups", "vote_stream", "collect_mode", "debug_novote", "join_parts", "join_hint",
             "verify_impl", "collect_unroll")
    crc0 = None
    for combo in combos:
        opts = dict(kv.split("=") for kv in combo.split(",") if kv)
        for k in names:
            mgr.set_option(k, int(opts.get(k, 1 if k == "join_impl" else 0)))
        acc = {}
        for it in range(8):
            res = mgr.search(qb)
            st, tm = res.stats()
            if it == 7:
                loops, cands = res.download()
            res.free()
            if it >= 3:
                for k, v in tm.items():
                    acc[k] = acc.get(k, 0.0) + v / 5
        import zlib
        crc = zlib.crc32(cands.tobytes(), zlib.crc32(loops.tobytes()))
        crc0 = crc0 if crc0 is not None else crc
        print(f"{combo:40s} vote {acc['vote_ms']:7.3f} probe {acc['probe_ms']:6.3f} topk {acc['topk_ms']:6.3f} "
              f"collect {acc['collect_ms']:6.3f} verify {acc['verify_ms']:6.3f} total {acc['total_ms']:7.3f}  "
              f"crc {'same' if crc == crc0 else 'DIFFERENT'}  M={st['M']}", flush=True)


if __name__ == "__main__":
    main()
