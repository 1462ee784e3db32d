"""Seeded synthetic "city" for the benchmark configs.

There is no network, so the reference's example dataset (MulRan / KITTI graphs)
is unavailable; every workload is generated here.  This module only produces
INPUTS (instance nodes per keyframe); it is not part of the hot path.

Node-level keyframes follow the shape of the reference's graph JSON
(R/include/Semantic_Graph.hpp:62-110): per keyframe a list of node labels
(3..11 after `node_map`, R/src/get_json.cpp:10-12) and float32 centres in the
sensor frame, plus a pose.
"""
import numpy as np

try:  # scipy is present in the image; used for the radius queries only
    from scipy.spatial import cKDTree
except Exception:  # pragma: no cover
    cKDTree = None

BASE_SEED = 0x5D7D0000

# node label -> (probability, z range) ; labels after node_map:
# 3 sidewalk, 4 other-ground, 5 building, 6 fence, 8 trunk, 9 terrain, 10 pole, 11 sign
_LABELS = np.array([3, 4, 5, 6, 8, 9, 10, 11], dtype=np.uint32)
_PROBS = np.array([0.02, 0.03, 0.15, 0.10, 0.25, 0.10, 0.25, 0.10])
_ZLO = np.array([-1.7, -1.7, 0.5, -1.0, -0.5, -1.7, 0.0, 0.5])
_ZHI = np.array([-1.5, -1.5, 6.0, 0.0, 1.5, -1.4, 3.0, 2.5])


def make_world(n_keyframes, seed, spacing=0.8, block=120.0, density=0.0055):
    """A Manhattan street grid, a loopy random-walk trajectory with one pose
    every `spacing` metres, and landmarks scattered over the grid area."""
    rng = np.random.default_rng(seed)
    length = n_keyframes * spacing
    G = max(3, int(np.sqrt(0.6 * length / (2.0 * block))) + 2)
    # random walk on intersections, no immediate U-turn
    pos = np.array([G // 2, G // 2])
    prev_dir = None
    dirs = np.array([[1, 0], [-1, 0], [0, 1], [0, -1]])
    pts = [pos * block]
    total = 0.0
    while total < length + block:
        cand = []
        for d in range(4):
            nxt = pos + dirs[d]
            if (nxt < 0).any() or (nxt >= G).any():
                continue
            if prev_dir is not None and (dirs[d] == -dirs[prev_dir]).all():
                continue
            cand.append(d)
        d = cand[rng.integers(len(cand))]
        pos = pos + dirs[d]
        prev_dir = d
        pts.append(pos * block)
        total += block
    pts = np.asarray(pts, dtype=np.float64)
    seg = np.diff(pts, axis=0)
    seglen = np.linalg.norm(seg, axis=1)
    cum = np.concatenate([[0.0], np.cumsum(seglen)])
    s = np.arange(n_keyframes) * spacing
    k = np.clip(np.searchsorted(cum, s, side="right") - 1, 0, len(seg) - 1)
    frac = (s - cum[k]) / seglen[k]
    xy = pts[k] + seg[k] * frac[:, None]
    yaw = np.arctan2(seg[k, 1], seg[k, 0])
    # lateral lane wobble so revisits are not pixel-identical
    xy = xy + rng.normal(0.0, 0.3, xy.shape)
    poses = np.column_stack([xy, yaw])
    # landmarks
    lo, hi = -100.0, (G - 1) * block + 100.0
    n_lm = int(density * (hi - lo) ** 2)
    lm_xy = rng.uniform(lo, hi, (n_lm, 2))
    cls = rng.choice(len(_LABELS), size=n_lm, p=_PROBS)
    lm_z = rng.uniform(_ZLO[cls], _ZHI[cls])
    lm = np.column_stack([lm_xy, lm_z])
    return dict(poses=poses, landmarks=lm, labels=_LABELS[cls], seed=seed, G=G, block=block)


def keyframes_at(world, poses, noise_seed, jitter=0.05, radius=80.0, kmin=10, kmax=200,
                 dropout=0.0):
    """Instance nodes seen from each pose: landmarks within `radius`, moved to
    the sensor frame, + N(0, jitter) noise; K clipped to [kmin, kmax] (nearest
    first).  Returns (xyz float32 [N,3], label uint32 [N], offsets int64 [n+1]).
    Frames with fewer than kmin nodes get the kmin nearest landmarks."""
    rng = np.random.default_rng(noise_seed)
    lm = world["landmarks"]
    tree = cKDTree(lm[:, :2])
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 3)
    n = poses.shape[0]
    lists = tree.query_ball_point(poses[:, :2], r=radius, return_sorted=True)
    lens = np.fromiter((len(l) for l in lists), dtype=np.int64, count=n)
    short = np.nonzero(lens < kmin)[0]
    if short.size:
        _, nn = tree.query(poses[short, :2], k=kmin)
        for j, f in enumerate(short):
            lists[f] = list(np.sort(nn[j]))
        lens[short] = kmin
    idx = np.concatenate([np.asarray(l, dtype=np.int64) for l in lists])
    frame = np.repeat(np.arange(n), lens)
    keep = np.ones(idx.shape[0], dtype=bool)
    if dropout > 0:
        keep = rng.random(idx.shape[0]) >= dropout
    d = lm[idx, :2] - poses[frame, :2]
    dist = np.hypot(d[:, 0], d[:, 1])
    # rank within frame by distance to enforce kmax / protect kmin from dropout
    order = np.lexsort((dist, frame))
    start = np.concatenate([[0], np.cumsum(lens)[:-1]])
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0]) - np.repeat(start, lens)
    keep = (keep | (rank < kmin)) & (rank < kmax)
    idx, frame, d = idx[keep], frame[keep], d[keep]
    c, s = np.cos(poses[frame, 2]), np.sin(poses[frame, 2])
    x = c * d[:, 0] + s * d[:, 1]
    y = -s * d[:, 0] + c * d[:, 1]
    z = lm[idx, 2]
    xyz = np.column_stack([x, y, z]) + rng.normal(0.0, jitter, (idx.shape[0], 3))
    lab = world["labels"][idx]
    # node order inside a frame mimics "ascending instance id" = class-major
    az = np.arctan2(xyz[:, 1], xyz[:, 0])
    order = np.lexsort((az, lab, frame))
    xyz, lab, frame = xyz[order], lab[order], frame[order]
    counts = np.bincount(frame, minlength=n)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return np.ascontiguousarray(xyz, dtype=np.float32), lab.astype(np.uint32), offsets


def make_queries(world, n_queries, seed, offset=2.0):
    """Query poses: DB places revisited with a +-offset shift and a random yaw.
    Returns (poses [n,3], gt_frame [n])."""
    rng = np.random.default_rng(seed)
    P = world["poses"]
    gt = rng.integers(0, P.shape[0], size=n_queries)
    q = P[gt].copy()
    q[:, :2] += rng.uniform(-offset, offset, (n_queries, 2))
    q[:, 2] = rng.uniform(-np.pi, np.pi, n_queries)
    return q, gt


def make_config(index, n_keyframes, n_queries, dropout=0.1):
    """World + DB keyframes + query keyframes for BASELINE.json config `index`."""
    seed = BASE_SEED + index
    w = make_world(n_keyframes, seed)
    db = keyframes_at(w, w["poses"], noise_seed=seed * 7 + 1)
    qposes, gt = make_queries(w, n_queries, seed * 7 + 2)
    qk = keyframes_at(w, qposes, noise_seed=seed * 7 + 3, dropout=dropout)
    return dict(world=w, db=db, queries=qk, qposes=qposes, gt=gt)
