"""CPU tests of the drop-in boundary: the shared library loads without a GPU,
exports every symbol include/sgtd_b200.h declares, and its host-side logic
(config, key packing, top-k merge) behaves like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest

from sgtd_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sgtd_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sgtd_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    assert capi.lib().sgtd_abi_version() == 2


def test_struct_layouts_match_header(tmp_path):
    """Compile the header as plain C and compare sizeof/offsetof with the binding."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sgtd_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(sgtd_config),sizeof(sgtd_desc),sizeof(sgtd_candidate),sizeof(sgtd_node),sizeof(sgtd_loop_result),"
                   "offsetof(sgtd_config,icp_threshold),offsetof(sgtd_candidate,R),offsetof(sgtd_candidate,inlier_off),"
                   "sizeof(sgtd_timings));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [ctypes.sizeof(capi.Config), capi.DESC_DTYPE.itemsize, capi.CAND_DTYPE.itemsize,
                   capi.NODE_DTYPE.itemsize, capi.LOOP_DTYPE.itemsize, capi.Config.icp_threshold.offset,
                   capi.CAND_DTYPE.fields["R"][1], capi.CAND_DTYPE.fields["inlier_off"][1],
                   ctypes.sizeof(capi.Timings)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.SgtdError) as e:
        capi.STDescManager()
    assert e.value.status == capi.E_CUDA


def test_config_defaults_and_yaml():
    c = capi.default_config()
    live = dict(descriptor_near_num=10, descriptor_min_len=0.5, descriptor_max_len=50.0, std_side_resolution=1.0,
                rough_dis_threshold=0.03, candidate_num=50, icp_threshold=0.4)
    for k, v in live.items():
        assert getattr(c, k) == v, k
    y = capi.config_from_yaml(os.path.join(ROOT, "tests", "golden", "sg_localization_std.yaml"))
    for name, _ in capi.Config._fields_:
        if name in ("stop_skip_enable", "plane_merge_dis_thre"):
            continue  # never read by read_parameters
        assert getattr(y, name) == getattr(c, name), name
    with pytest.raises(capi.SgtdError):
        capi.config_from_yaml("/nonexistent.yaml")


def test_db_key_packing(oracle_lib):
    rng = np.random.default_rng(3)
    d = np.zeros(200, capi.DESC_DTYPE)
    d["side"] = np.sort(rng.uniform(0.5, 50, (200, 3)), axis=1)
    d["side"][:5] = [[0.5, 0.5, 0.5], [1.49999, 1.5, 2.5], [49.5, 49.6, 50.0], [0.99, 1.0, 1.01], [7.5, 7.5, 7.5]]
    d["lab"] = rng.integers(0, 20, (200, 3))
    ok = oracle_lib.Oracle.db_keys(d.view(oracle_lib.DESC_DTYPE))
    for i in range(200):
        key = capi.db_key(d[i])
        assert (key >> 44, (key >> 28) & 0xFFFF, (key >> 12) & 0xFFFF, key & 0xFFF) == tuple(int(x) for x in ok[i])


def test_merge_topk_host():
    rng = np.random.default_rng(4)
    k = 50
    for nl in (1, 2, 8):
        votes = np.zeros((nl, k), np.int32)
        frames = np.full((nl, k), -1, np.int32)
        allv = {}
        for r in range(nl):
            n = rng.integers(0, k + 1)
            f = rng.choice(1000, size=n, replace=False) + 1000 * r
            v = rng.integers(5, 12, size=n)  # many ties
            o = np.lexsort((f, -v))
            votes[r, :n], frames[r, :n] = v[o], f[o]
            allv.update(dict(zip(f.tolist(), v.tolist())))
        ov, of = capi.merge_topk_host(votes, frames, k)
        exp = sorted(allv.items(), key=lambda t: (-t[1], t[0]))[:k]
        n = len(exp)
        assert of[:n].tolist() == [e[0] for e in exp] and ov[:n].tolist() == [e[1] for e in exp]
        assert (ov[n:] == 0).all() and (of[n:] == -1).all()
