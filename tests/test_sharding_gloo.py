"""N>1 path on CPU (gloo, world_size 2): keyframe-range shards vote
independently, per-shard top-k lists are all-gathered and merged with the
library's deterministic merge (the same routine the GPU merge kernel runs).
The merged ranking must equal the unsharded oracle's candidate list."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

N_KF, N_Q, K = 96, 3, 50


def _local_topk(votes, lo, hi, k):
    f = np.arange(lo, hi)
    v = votes[lo:hi]
    keep = v >= 5
    f, v = f[keep], v[keep]
    o = np.lexsort((f, -v))[:k]
    ov = np.zeros(k, np.int32)
    of = np.full(k, -1, np.int32)
    ov[:o.size], of[:o.size] = v[o], f[o]
    return ov, of


def _worker(rank, world, port, q):
    import torch
    from oracle import orc
    from sgtd_b200 import capi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.make_config(0, N_KF, N_Q)
    xyz, lab, off = cfg["db"]
    qx, ql, qo = cfg["queries"]
    fpr = (N_KF + world - 1) // world
    lo, hi = rank * fpr, min(N_KF, (rank + 1) * fpr)
    shard = orc.Oracle()
    full = orc.Oracle() if rank == 0 else None
    for f in range(N_KF):
        d = shard.build(xyz[off[f]:off[f + 1]], lab[off[f]:off[f + 1]])
        shard.add(d if lo <= f < hi else d[:0])   # same frame ids, only owned keyframes stored
        if full is not None:
            full.add(full.build(xyz[off[f]:off[f + 1]], lab[off[f]:off[f + 1]]))
    ok = True
    for qi in range(N_Q):
        qd = shard.build(qx[qo[qi]:qo[qi + 1]], ql[qo[qi]:qo[qi + 1]])
        r = shard.search(qd)
        ov, of = _local_topk(r["votes"], lo, hi, K)
        # shard-local selector must agree with the local top-k rule
        assert list(r["cands"]["frame"]) == [x for x in of if x >= 0]
        gv = [torch.zeros(K, dtype=torch.int32) for _ in range(world)]
        gf = [torch.zeros(K, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(gv, torch.from_numpy(ov))
        dist.all_gather(gf, torch.from_numpy(of))
        mv, mf = capi.merge_topk_host(np.stack([t.numpy() for t in gv]), np.stack([t.numpy() for t in gf]), K)
        if rank == 0:
            ref = full.search(full.build(qx[qo[qi]:qo[qi + 1]], ql[qo[qi]:qo[qi + 1]]))
            n = ref["n"]
            ok &= mf[:n].tolist() == ref["cands"]["frame"].tolist() and mv[:n].tolist() == ref["cands"]["votes"].tolist()
            ok &= bool((mf[n:] == -1).all())
            # the owner's match list of a merged candidate equals the unsharded one
            for c, fr in enumerate(ref["cands"]["frame"][:5]):
                if lo <= fr < hi:
                    j = list(r["cands"]["frame"]).index(fr)
                    a, b = r["cands"][j], ref["cands"][c]
                    ok &= a["nmatch"] == b["nmatch"] and a["score"] == b["score"]
    if rank == 0:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_shards_merge_equals_unsharded(oracle_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
