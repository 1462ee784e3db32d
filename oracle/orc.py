"""ctypes wrapper of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by sgtd_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")

DESC_DTYPE = np.dtype(
    [("side", "<f8", 3), ("vert", "<f4", 9), ("frame", "<u4"), ("lab", "u1", 3),
     ("pad", "u1"), ("anchor", "<u2"), ("m", "u1"), ("n", "u1")], align=False)
assert DESC_DTYPE.itemsize == 72

CAND_DTYPE = np.dtype(
    [("frame", "<i4"), ("votes", "<i4"), ("nmatch", "<i4"), ("score", "<i4"),
     ("match_off", "<i4"), ("inlier_off", "<i4"), ("ninlier", "<i4"),
     ("best_hyp", "<i4"), ("R", "<f8", 9), ("t", "<f8", 3)], align=False)
assert CAND_DTYPE.itemsize == 128


class OrcConfig(C.Structure):
    _fields_ = [("descriptor_near_num", C.c_int32), ("candidate_num", C.c_int32),
                ("descriptor_min_len", C.c_double), ("descriptor_max_len", C.c_double),
                ("std_side_resolution", C.c_double), ("rough_dis_threshold", C.c_double),
                ("icp_threshold", C.c_double)]


class OrcStats(C.Structure):
    _fields_ = [("Q", C.c_int64), ("P", C.c_int64), ("Pfound", C.c_int64),
                ("E", C.c_int64), ("M", C.c_int64)]


def build(force=False):
    if force or not os.path.exists(_LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
            for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_current_frame_id.restype = C.c_uint32
        L.orc_current_frame_id.argtypes = [C.c_void_p]
        L.orc_db_size.restype = C.c_int64
        L.orc_db_size.argtypes = [C.c_void_p]
        L.orc_build.restype = C.c_int64
        L.orc_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
        L.orc_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_db_key.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_build_add_many.restype = C.c_int64
        L.orc_build_add_many.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.orc_search.restype = C.c_int32
        L.orc_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(OrcStats), C.c_int32]
        L.orc_triangle_solver.argtypes = [C.c_void_p] * 4
        L.orc_jacobi_svd3.argtypes = [C.c_void_p] * 4
        L.orc_dcvc.restype = C.c_int32
        L.orc_dcvc.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32,
                               C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_extract_instances.restype = C.c_int32
        L.orc_extract_instances.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_extract_instances_submap.restype = C.c_int32
        L.orc_extract_instances_submap.argtypes = L.orc_extract_instances.argtypes
        L.orc_submap_aggregate.restype = C.c_int64
        L.orc_submap_aggregate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                           C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_gicp_align.restype = C.c_int32
        L.orc_gicp_align.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


DEFAULT_CFG = dict(descriptor_near_num=10, candidate_num=50, descriptor_min_len=0.5,
                   descriptor_max_len=50.0, std_side_resolution=1.0,
                   rough_dis_threshold=0.03, icp_threshold=0.4)


class Oracle:
    """Mirror of STDescManager (reference: R/include/desc/STDesc.h:342-440)."""

    def __init__(self, **cfg):
        c = dict(DEFAULT_CFG)
        c.update(cfg)
        self.cfg = c
        self._h = lib().orc_create(C.byref(OrcConfig(**c)))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    @property
    def current_frame_id(self):
        return lib().orc_current_frame_id(self._h)

    @property
    def db_size(self):
        return lib().orc_db_size(self._h)

    def build(self, xyz, label):
        """BuildSingleScanSTD -> structured array of descriptors."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        label = np.ascontiguousarray(label, dtype=np.uint32)
        K = xyz.shape[0]
        near = self.cfg["descriptor_near_num"]
        cap = max((near - 1) * (near - 2) // 2 * K, 1)  # all (m, n) pairs of every anchor
        out = np.zeros(cap, dtype=DESC_DTYPE)
        n = lib().orc_build(self._h, _p(xyz), _p(label), K, _p(out), cap)
        if n < 0:
            raise ValueError("too few nodes (K < descriptor_near_num)")
        return out[:n].copy()

    def add(self, descs):
        descs = np.ascontiguousarray(descs, dtype=DESC_DTYPE)
        lib().orc_add(self._h, _p(descs), descs.shape[0])

    def build_add_many(self, xyz, label, off, nthreads=None):
        """map phase for many keyframes: build (parallel) + add (in order); same DB as per-frame calls"""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        label = np.ascontiguousarray(label, dtype=np.uint32)
        off = np.ascontiguousarray(off, dtype=np.int64)
        return lib().orc_build_add_many(self._h, _p(xyz), _p(label), _p(off), off.shape[0] - 1,
                                        nthreads or os.cpu_count() or 1)

    @staticmethod
    def db_keys(descs):
        descs = np.ascontiguousarray(descs, dtype=DESC_DTYPE)
        out = np.zeros((descs.shape[0], 4), dtype=np.int32)
        for i in range(descs.shape[0]):
            lib().orc_db_key(_p(descs[i:i + 1]), _p(out[i]))
        return out

    def search(self, q, nthreads=1, cap_match=None, want_votes=True):
        """SearchLoop.  Returns dict(cands, m_q, m_cell, m_g, inl, votes, best, stats)."""
        q = np.ascontiguousarray(q, dtype=DESC_DTYPE)
        ncap = self.cfg["candidate_num"]
        cands = np.zeros(ncap, dtype=CAND_DTYPE)
        if cap_match is None:
            cap_match = 1 << 22
        F = self.current_frame_id
        while True:
            m_q = np.zeros(cap_match, np.int32)
            m_cell = np.zeros(cap_match, np.uint8)
            m_g = np.zeros(cap_match, np.uint32)
            inl = np.zeros(cap_match, np.int32)
            votes = np.zeros(max(F, 1), np.int32) if want_votes else None
            best = np.zeros(2, np.float64)
            st = OrcStats()
            n = lib().orc_search(self._h, _p(q), q.shape[0], _p(cands), ncap, _p(m_q), _p(m_cell),
                                 _p(m_g), _p(inl), cap_match, _p(votes), F, _p(best),
                                 C.byref(st), nthreads)
            if n == -2:
                cap_match *= 4
                continue
            break
        stats = {k: getattr(st, k) for k in ("Q", "P", "Pfound", "E", "M")}
        if n < 0:
            return dict(n=n, cands=cands[:0], m_q=m_q[:0], m_cell=m_cell[:0], m_g=m_g[:0],
                        inl=inl[:0], votes=votes, best=(-1, 0.0), stats=stats)
        cands = cands[:n]
        nm = int(cands["nmatch"].sum())
        ni = int(cands["ninlier"].sum())
        return dict(n=n, cands=cands, m_q=m_q[:nm], m_cell=m_cell[:nm], m_g=m_g[:nm],
                    inl=inl[:ni], votes=votes, best=(int(best[0]), float(best[1])), stats=stats)


def triangle_solver(src, ref):
    src = np.ascontiguousarray(src, dtype=DESC_DTYPE).reshape(1)
    ref = np.ascontiguousarray(ref, dtype=DESC_DTYPE).reshape(1)
    R = np.zeros(9)
    t = np.zeros(3)
    lib().orc_triangle_solver(_p(src), _p(ref), _p(R), _p(t))
    return R.reshape(3, 3), t


def jacobi_svd3(A):
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(9)
    U = np.zeros(9)
    s = np.zeros(3)
    V = np.zeros(9)
    lib().orc_jacobi_svd3(_p(A), _p(U), _p(s), _p(V))
    return U.reshape(3, 3), s, V.reshape(3, 3)


def dcvc(xyz, startR=0.35, deltaR=0.0004, deltaP=1.2, deltaA=1.2, minSeg=300):
    """clusterManager::segmentPointCloud on one class cloud.
    Returns (label_info[n], cluster_of[n], n_clusters, (width, height, polarNum))."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    n = xyz.shape[0]
    lab = np.full(n, -1, np.int32)
    cl = np.full(n, -1, np.int32)
    grid = np.zeros(3, np.int32)
    nc = lib().orc_dcvc(_p(xyz), n, startR, deltaR, deltaP, deltaA, minSeg, _p(lab), _p(cl), _p(grid))
    return lab, cl, int(nc), tuple(int(x) for x in grid)


def submap_aggregate(points, labels, poses12, j, b2o=None, radius=15.0):
    """the point-gathering part of local_map_creation (R/src/local_map.cpp:213-328), literally.
    -> (points [m,4], labels [m], scans used)"""
    points = np.ascontiguousarray(points, np.float32).reshape(-1, 4)
    labels = np.ascontiguousarray(labels, np.uint32)
    poses = np.ascontiguousarray(poses12, np.float32).reshape(-1, 12)
    b = np.eye(4, dtype=np.float32) if b2o is None else np.ascontiguousarray(b2o, np.float32).reshape(4, 4)
    cap = points.shape[0] * poses.shape[0]
    op = np.zeros((cap, 4), np.float32)
    ol = np.zeros(cap, np.uint32)
    used = C.c_int32(0)
    m = lib().orc_submap_aggregate(_p(points), _p(labels), points.shape[0], _p(poses), poses.shape[0], j, _p(b), radius,
                                   _p(op), _p(ol), cap, C.byref(used))
    return op[:m].copy(), ol[:m].copy(), used.value


def extract_instances(points, labels, submap=False):
    """gen_labels + gen_graphs nodes (submap=True: the class tables of local_map_creation).
    points [n,4] float32, labels [n] uint32.
    Returns dict(point_instance[n], node_xyz[k,3], node_label[k], node_inst[k], n_instances)."""
    points = np.ascontiguousarray(points, np.float32).reshape(-1, 4)
    labels = np.ascontiguousarray(labels, np.uint32)
    n = points.shape[0]
    cap = 1 << 16
    pi = np.full(n, -1, np.int32)
    nx = np.zeros((cap, 3), np.float32)
    nl = np.zeros(cap, np.uint32)
    ni = np.zeros(cap, np.int32)
    nn = C.c_int32(0)
    ninst = C.c_int32(0)
    fn = lib().orc_extract_instances_submap if submap else lib().orc_extract_instances
    rc = fn(_p(points), _p(labels), n, _p(pi), _p(nx), _p(nl), _p(ni), cap, C.byref(nn), C.byref(ninst))
    if rc:
        raise RuntimeError("orc_extract_instances: capacity")
    k = nn.value
    return dict(point_instance=pi, node_xyz=nx[:k].copy(), node_label=nl[:k].copy(), node_inst=ni[:k].copy(),
                n_instances=ninst.value)


def gicp_align(source, target, init12=None, k=20, max_iterations=10, rot_eps=2e-3, trans_eps=5e-4):
    """fast_gicp::FastGICP::align as the node calls it -> (final 4x4, fitness, iterations, converged)."""
    source = np.ascontiguousarray(source, np.float32).reshape(-1, 3)
    target = np.ascontiguousarray(target, np.float32).reshape(-1, 3)
    init = np.eye(4)[:3].reshape(12).copy() if init12 is None else np.ascontiguousarray(init12, np.float64).reshape(12)
    fin = np.zeros(16)
    fit = C.c_double(0)
    it, cv = C.c_int32(0), C.c_int32(0)
    rc = lib().orc_gicp_align(_p(source), source.shape[0], _p(target), target.shape[0], _p(init), k, max_iterations,
                              rot_eps, trans_eps, _p(fin), C.byref(fit), C.byref(it), C.byref(cv))
    if rc:
        raise ValueError("orc_gicp_align: clouds smaller than k")
    return fin.reshape(4, 4), fit.value, it.value, bool(cv.value)
