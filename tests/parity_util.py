"""Shared comparison helpers of the GPU parity tests (CUDA path through the C ABI vs the CPU oracle /
the reference build)."""
import numpy as np

POSE_T_TOL = 0.01         # 1 cm            (north_star)
POSE_R_TOL_DEG = 0.01     # 0.01 degree


def rot_angle_deg(Ra, Rb):
    d = Ra.reshape(3, 3).T @ Rb.reshape(3, 3)
    return np.degrees(np.arccos(np.clip((np.trace(d) - 1) / 2, -1, 1)))


def compare_query(res, loops, cands, q, r, n_frames=None, votes=True, cells=True, max_lists=None):
    """Query q of a GPU search result against the oracle's (or the reference build's) result r of the
    same query: per-keyframe votes, candidate ranking, match lists, hypothesis choice, inlier lists and
    scores bit-exact; pose within 1 cm / 0.01 deg (and reported if not bit-equal).
    votes / cells: r carries per-keyframe votes / probe ordinals (the reference build does not).
    Returns the number of candidates whose pose was bit-equal."""
    if votes:
        assert (res.votes(q, n_frames) == r["votes"]).all(), f"votes of query {q}"
    n = r["n"]
    assert loops["ncand"][q] == n, (q, loops["ncand"][q], n)
    oc, gc = r["cands"], cands[q, :n]
    assert (gc["frame"] == oc["frame"]).all(), f"ranking of query {q}"
    assert (gc["nmatch"] == oc["nmatch"]).all() and (gc["votes"] == oc["nmatch"]).all()
    assert (cands["frame"][q, n:] == -1).all()
    assert (gc["score"] == oc["score"]).all(), f"scores of query {q}"
    assert (gc["ninlier"][oc["score"] >= 0] == oc["ninlier"][oc["score"] >= 0]).all()
    if "best_hyp" in oc.dtype.names and (oc["best_hyp"] != -1).any():
        assert (gc["best_hyp"] == oc["best_hyp"]).all()
    exact = 0
    moff = np.concatenate([[0], np.cumsum(oc["nmatch"])])
    ioff = np.concatenate([[0], np.cumsum(np.maximum(oc["ninlier"], 0))])
    for c in range(n if max_lists is None else min(n, max_lists)):
        m_q, m_cell, m_g = res.matches(q, c, int(gc["nmatch"][c]))
        s = slice(moff[c], moff[c + 1])
        assert (m_q == r["m_q"][s]).all() and (m_g == r["m_g"][s]).all(), f"match list {q}/{c}"
        if cells:
            assert (m_cell == r["m_cell"][s]).all()
        if oc["score"][c] >= 0:
            inl = res.inliers(q, c, int(gc["ninlier"][c]))
            assert (inl == r["inl"][ioff[c]:ioff[c + 1]]).all(), f"inliers {q}/{c}"
            assert np.abs(gc["t"][c] - oc["t"][c]).max() <= POSE_T_TOL
            assert rot_angle_deg(gc["R"][c], oc["R"][c]) <= POSE_R_TOL_DEG
            exact += gc["R"][c].tobytes() == oc["R"][c].tobytes() and gc["t"][c].tobytes() == oc["t"][c].tobytes()
    assert loops["frame"][q] == r["best"][0] and loops["score"][q] == r["best"][1]
    return exact
