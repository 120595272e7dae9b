"""hpgmg_b200 -- B200-native HPGMG-FV (fv4 full-multigrid F-cycle) behind the reference's C API.

The product is hpgmg_b200/lib/libhpgmg_b200.so (C host code + hand-written sm_100a kernels, built
from hpgmg_b200/csrc by __graft_entry__.build()); `hpgmg_b200.api` is the thin ctypes host layer.
"""
from . import api  # noqa: F401

__all__ = ["api"]
