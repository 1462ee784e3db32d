"""The C++ facade (include/sgtd/STDesc.h) compiles as plain C++17 against the C ABI
(CPU check) and, on the GPU box, reproduces the oracle when driven with the
reference node's call sequence (tests/cpp/facade_node.cpp)."""
import os
import subprocess

import numpy as np
import pytest

from sgtd_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "sgtd_b200", "lib")


def _compile(tmp_path, eigen_like=False):
    """eigen_like: resolve <Eigen/Core>, <pcl/...>, <ros/ros.h> to the Eigen/PCL/ROS-like headers of
    oracle/shim (the ones the reference's own sources compile against in oracle/_ref), so that the facade
    takes its SGTD_HAVE_EIGEN / _PCL / _ROS branches and the node's Eigen expressions (block<3,3>, <<,
    cast<float>, products) run on the facade's STDesc / LOOP_RESULT types.  Otherwise: its POD stand-ins."""
    exe = str(tmp_path / ("facade_node_eigen" if eigen_like else "facade_node"))
    extra = ["-I", os.path.join(ROOT, "oracle", "shim")] if eigen_like else []
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include")] + extra +
                          [os.path.join(ROOT, "tests", "cpp", "facade_node.cpp"), "-o", exe,
                           "-L", LIBDIR, "-lsgtd_b200", f"-Wl,-rpath,{LIBDIR}"])
    return exe


@pytest.mark.parametrize("eigen_like", [False, True])
def test_facade_compiles_and_fails_loudly_without_gpu(tmp_path, eigen_like):
    capi.lib()
    exe = _compile(tmp_path, eigen_like)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    bin_path = tmp_path / "c.bin"
    bin_path.write_bytes(np.zeros(2, np.int32).tobytes())
    p = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "sg_localization_std.yaml"), str(bin_path)],
                       capture_output=True, text=True)
    assert p.returncode != 0 and "SGTD_E_CUDA" in (p.stderr + p.stdout)   # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("eigen_like", [False, True])
def test_facade_reproduces_oracle(tmp_path, oracle_lib, eigen_like):
    exe = _compile(tmp_path, eigen_like)
    cfg = synth.make_config(0, 40, 2)
    xyz, lab, off = cfg["db"]
    qx, ql, qo = cfg["queries"]
    buf = [np.array([off.shape[0] - 1, qo.shape[0] - 1], np.int32).tobytes()]
    for (X, L, O) in ((xyz, lab, off), (qx, ql, qo)):
        for s in range(O.shape[0] - 1):
            n = capi.make_nodes(X[O[s]:O[s + 1]], L[O[s]:O[s + 1]])
            buf.append(np.array([n.shape[0]], np.int32).tobytes() + n.tobytes())
    (tmp_path / "c.bin").write_bytes(b"".join(buf))
    out = subprocess.check_output([exe, os.path.join(ROOT, "tests", "golden", "sg_localization_std.yaml"),
                                   str(tmp_path / "c.bin")], text=True).splitlines()
    assert out[0] == "frames 40"
    o = oracle_lib.Oracle()
    for f in range(40):
        o.add(o.build(xyz[off[f]:off[f + 1]], lab[off[f]:off[f + 1]]))
    for q in range(2):
        qd = o.build(qx[qo[q]:qo[q + 1]], ql[qo[q]:qo[q + 1]])
        r = o.search(qd)
        tok = out[1 + q].split()
        assert int(tok[3]) == len(qd)
        assert int(tok[5]) == r["best"][0] and float(tok[6]) == r["best"][1]
        assert int(tok[10]) == r["n"]
        if eigen_like:   # the node's own Eigen pose arithmetic agrees with sgtd_localization_check
            assert tok[tok.index("eigen_check") + 1] == "1"
        cands = tok[tok.index("cands") + 1:]
        exp = [f"{c['frame']}:{c['score']}:{max(c['ninlier'], 0) if c['score'] > 0 else 0}" for c in r["cands"]]
        assert cands == exp
        best = [c for c in r["cands"] if c["frame"] == r["best"][0] and c["score"] == r["best"][1]][0]
        assert int(tok[8]) == best["ninlier"]
        t = np.array([float(x) for x in tok[12:15]])
        R = np.array([float(x) for x in tok[16:25]])
        assert np.abs(t - best["t"]).max() < 1e-6 and np.abs(R - best["R"]).max() < 1e-9
