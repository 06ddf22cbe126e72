"""The C-ABI library loads without a GPU and exports every symbol include/dpvo_b200.h declares."""
import ctypes
import os
import re

import dpvo_b200

HDR = os.path.join(dpvo_b200.INCLUDE_DIR, "dpvo_b200.h")


def _declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpvo_[a-zA-Z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = dpvo_b200.library()
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in dpvo_b200.h but not exported: %s" % missing


def test_version_and_error_text():
    lib = dpvo_b200.library()
    assert b"sm_100a" in lib.dpvo_version()
    assert isinstance(lib.dpvo_last_error(), bytes)
    assert lib.dpvo_launch_count() >= 0


def test_argument_validation_needs_no_gpu():
    """bad arguments are rejected on the host before any CUDA call"""
    lib = dpvo_b200.library()
    rc = lib.dpvo_corr_forward(None, None, None, None, None, None, None, None, ctypes.c_int64(1),
                               1, 1, 4, 128, 3, 1, 1, 8, 8, 3, None)
    assert rc == 1 and b"null" in lib.dpvo_last_error()
    lib.dpvo_ba_workspace_bytes.restype = ctypes.c_int64
    lib.dpvo_ba_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int]
    assert lib.dpvo_ba_workspace_bytes(47712, 10) > 47712 * 60 * 4


def test_shims_import_and_refuse_cpu_tensors():
    import pytest
    import torch
    cc, cb, lb, ex = dpvo_b200.extensions()
    assert {"forward", "backward", "patchify_forward", "patchify_backward"} <= set(dir(cc))
    assert {"forward", "neighbors", "reproject", "solve_system"} <= set(dir(cb))
    assert {"expm", "expm_backward", "logm", "logm_backward", "inv", "inv_backward", "mul", "mul_backward", "adj",
            "adj_backward", "adjT", "adjT_backward", "act", "act_backward", "act4", "act4_backward", "as_matrix",
            "projector", "Jinv"} <= set(dir(lb))
    with pytest.raises(RuntimeError, match="no CPU path"):
        lb.expm(3, torch.zeros(2, 6))
    with pytest.raises(RuntimeError, match="no CPU path"):
        cb.neighbors(torch.zeros(3, dtype=torch.long), torch.zeros(3, dtype=torch.long))
