"""Host logic of the multi-GPU bench layouts (CPU): S keyframe-range shards x R replica groups cover every
keyframe exactly once per group and every query exactly once over the groups."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


@pytest.mark.parametrize("world,shards", [(1, 0), (2, 0), (2, 1), (4, 2), (8, 0), (8, 2), (8, 1), (8, 4)])
@pytest.mark.parametrize("nkf,nq", [(100000, 1024), (10001, 256), (7, 8)])
def test_layout_covers_frames_and_queries(world, shards, nkf, nq):
    plans = [bench.plan_layout(r, world, shards, nkf, nq) for r in range(world)]
    S, R = plans[0]["S"], plans[0]["R"]
    assert S * R == world and S == (shards or world)
    for g in range(R):
        grp = [p for p in plans if p["group"] == g]
        assert [p["shard_rank"] for p in grp] == list(range(S))
        # the group's shards tile [0, nkf) in order, with the range arithmetic sgtd_shard_init uses
        pos = 0
        for p in grp:
            assert p["frame_lo"] == min(pos, p["frame_lo"]) and p["frame_lo"] <= p["frame_hi"] <= nkf
            if p["frame_hi"] > p["frame_lo"]:
                assert p["frame_lo"] == pos == p["shard_rank"] * p["frames_per_rank"] or S == 1
                pos = p["frame_hi"]
        assert pos == nkf
        # every rank of a group serves the same query slice
        assert len({(p["query_lo"], p["query_hi"]) for p in grp}) == 1
    slices = sorted({(p["query_lo"], p["query_hi"]) for p in plans})
    assert slices[0][0] == 0 and slices[-1][1] == nq
    assert all(a[1] == b[0] for a, b in zip(slices, slices[1:]))


def test_layout_rejects_bad_shard_counts():
    with pytest.raises(ValueError):
        bench.plan_layout(0, 8, 3, 1000, 64)
