"""ORACLE helper (test infrastructure): import the reference's Python package `dpvo` from
/root/reference in the build container, with oracle/shims/ standing in for the modules that cannot
exist here (lietorch_backends, cuda_corr, cuda_ba: native + need a GPU; torch_scatter, pypose,
matplotlib, yacs: not installed).  Used only to generate golden vectors and to pin the restatements;
/root/reference does not exist on the GPU box."""
import contextlib
import os
import sys

REF_ROOT = "/root/reference"
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_NAMES = ("dpvo", "lietorch_backends", "cuda_corr", "cuda_ba", "torch_scatter", "pypose", "matplotlib", "yacs")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "dpvo"))


@contextlib.contextmanager
def reference_modules():
    """Context manager: inside it `import dpvo.ba`, `dpvo.projective_ops`, `dpvo.net` ... resolve to
    the reference's files.  sys.modules / sys.path are restored on exit."""
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REF_ROOT)
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _NAMES}
    for k in list(saved_mods):
        del sys.modules[k]
    sys.path[:0] = [SHIMS, REF_ROOT]
    try:
        yield
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _NAMES]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
