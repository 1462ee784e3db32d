"""GICP refinement of the verified candidates (the reference's final method,
R/src/semantic_graph_localization.cpp:651-721 + R/include/fast_gicp/gicp/impl/*.hpp).

CPU part: the oracle restatement (oracle/gicp_oracle.cpp) recovers known rigid motions and its
pieces agree with independent numpy re-derivations.  GPU part: the CUDA path (sgtd_gicp_align /
sgtd_gicp_refine_candidates through the C ABI) against the oracle on the same seeded clouds.
Floating point, tolerance stated here: final transform within 1e-5 (entries of the float matrix
the reference returns; observed ~1e-7), fitness within 1e-5 relative, same iteration count."""
import numpy as np
import pytest

T_TOL = 1e-5
FIT_RTOL = 1e-5


def scene(rng, n):
    g = np.column_stack([rng.uniform(-20, 20, n // 2), rng.uniform(-20, 20, n // 2), rng.normal(0, 0.02, n // 2)])
    w1 = np.column_stack([rng.uniform(-20, 20, n // 4), 8.0 + rng.normal(0, 0.02, n // 4), rng.uniform(0, 5, n // 4)])
    w2 = np.column_stack([-6.0 + rng.normal(0, 0.02, n // 4), rng.uniform(-20, 20, n // 4), rng.uniform(0, 4, n // 4)])
    return np.concatenate([g, w1, w2]).astype(np.float32)


def motion(rng, ang_deg, trans):
    a = np.radians(ang_deg)
    axis = rng.normal(size=3) * [0.15, 0.15, 1.0]
    axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K
    return R, np.asarray(trans, float)


def make_pair(seed, ang=3.0, trans=(0.4, -0.3, 0.05), n_t=4000, n_s=1500):
    rng = np.random.default_rng(seed)
    tgt = scene(rng, n_t)
    R, t = motion(rng, ang, trans)
    src = ((scene(rng, n_s) - t) @ R).astype(np.float32)      # p_target = R p_source + t
    return src, tgt, R, t


def test_oracle_recovers_known_motion(oracle_lib):
    for seed in range(4):
        src, tgt, R, t = make_pair(seed, ang=2.0 + seed, trans=(0.3 + 0.1 * seed, -0.2, 0.05))
        F, fit, it, conv = oracle_lib.gicp_align(src, tgt, k=20, max_iterations=20)
        assert conv and np.abs(F[:3, :3] - R).max() < 5e-3 and np.abs(F[:3, 3] - t).max() < 5e-2
        assert fit < 0.3
        # the init transform is applied to the source first: starting at the solution needs (almost) no motion
        init = np.column_stack([R, t]).reshape(12)
        F2, fit2, it2, conv2 = oracle_lib.gicp_align(src, tgt, init12=init, k=20, max_iterations=20)
        assert conv2 and np.abs(F2[:3, :3] - np.eye(3)).max() < 5e-3 and np.abs(F2[:3, 3]).max() < 5e-2
        assert abs(fit2 - fit) < 0.05


def test_oracle_fitness_is_mean_squared_nn_distance(oracle_lib):
    from scipy.spatial import cKDTree
    src, tgt, R, t = make_pair(11)
    F, fit, _, _ = oracle_lib.gicp_align(src, tgt, k=20, max_iterations=0)      # no iteration: identity
    assert np.array_equal(F, np.eye(4))
    d, _ = cKDTree(tgt.astype(np.float64)).query(src.astype(np.float64))
    assert abs(fit - np.mean(d ** 2)) < 1e-5 * max(1.0, fit)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,k,use_init", [(0, 20, False), (1, 20, True), (2, 10, False), (3, 40, True), (4, 20, False)])
def test_gpu_gicp_align_matches_oracle(oracle_lib, seed, k, use_init):
    from sgtd_b200 import capi
    src, tgt, R, t = make_pair(100 + seed, ang=1.5 + seed, trans=(0.5, -0.2 - 0.1 * seed, 0.03))
    init = None
    if use_init:   # a rough candidate pose, like the loop transform of stage 4
        Ri, ti = motion(np.random.default_rng(seed), 1.0 + seed, (0.45, -0.25, 0.0))
        init = np.column_stack([Ri, ti]).reshape(12)
    mgr = capi.STDescManager(device=0)
    prm = mgr.gicp_params(num_neighbors=k, max_iterations=12)
    F, fit, it, conv = mgr.gicp_align(src, tgt, init, prm)
    Fo, fito, ito, convo = oracle_lib.gicp_align(src, tgt, init12=init, k=k, max_iterations=12)
    assert it == ito and conv == convo
    assert np.abs(F - Fo[:3]).max() <= T_TOL
    assert abs(fit - fito) <= FIT_RTOL * max(1.0, abs(fito))
    # and the answer is the right one
    full = np.vstack([F, [0, 0, 0, 1]]) @ (np.vstack([init.reshape(3, 4), [0, 0, 0, 1]]) if use_init else np.eye(4))
    assert np.abs(full[:3, :3] - R).max() < 1e-2 and np.abs(full[:3, 3] - t).max() < 0.1
    mgr.close()


@pytest.mark.gpu
def test_gpu_multi_candidate_refinement_is_the_nodes_loop(oracle_lib):
    """visit candidates by fitness descending; lowest GICP fitness below 100 wins, the first below
    best_fitness ends the search (R/src/semantic_graph_localization.cpp:603, 671-721)."""
    from sgtd_b200 import capi
    rng = np.random.default_rng(7)
    src, tgt_good, R, t = make_pair(200)
    tgt_other = scene(np.random.default_rng(999), 4000) * np.float32(1.7) + np.float32(9.0)   # a different place
    mgr = capi.STDescManager(device=0)
    c = np.zeros(3, capi.CAND_DTYPE)
    c["frame"] = [5, 9, 2]
    c["score"] = [40, 60, 10]
    for i in range(3):
        Ri, ti = motion(rng, 1.0, (0.4, -0.3, 0.0))
        c["R"][i] = Ri.reshape(9)
        c["t"][i] = ti
    targets = [tgt_good, tgt_other, tgt_good]
    order = np.array([1, 0, 2], np.int32)                       # match_fitness descending
    for best_fitness in (0.5, 1e-9):
        prm = mgr.gicp_params(max_iterations=12, best_fitness=best_fitness)
        chosen, T, fit, nal = mgr.gicp_refine_candidates(src, targets, c, order, prm)
        exp_choice, exp_fit, exp_T, exp_n, bit = -1, 100.0, np.eye(4)[:3], 0, 100.0
        for ci in order:
            init = np.column_stack([c["R"][ci].reshape(3, 3), c["t"][ci]]).reshape(12)
            Fo, fo, _, _ = oracle_lib.gicp_align(src, targets[ci], init12=init, k=20, max_iterations=12)
            exp_n += 1
            if fo < bit:
                bit, exp_choice, exp_fit, exp_T = fo, int(ci), fo, Fo[:3]
            if fo < best_fitness:
                exp_choice, exp_fit, exp_T = int(ci), fo, Fo[:3]
                break
        assert chosen == exp_choice and nal == exp_n
        assert abs(fit - exp_fit) <= FIT_RTOL * max(1.0, exp_fit) and np.abs(T - exp_T).max() <= T_TOL
    assert chosen == 0          # the wrong place (visited first) loses to the right one
    mgr.close()
