"""sgtd_b200 -- B200 (sm_100a) implementation of SGTD's one-shot localization
hot path behind the reference's descriptor-manager surface.

The product is sgtd_b200/lib/libsgtd_b200.so (C ABI: include/sgtd_b200.h);
`capi` is the ctypes binding used by tests and bench.py, `synth` generates the
synthetic workloads.  Importing `capi.lib()` without the built library raises.
"""
from . import capi, synth  # noqa: F401
from .capi import STDescManager, default_config  # noqa: F401
