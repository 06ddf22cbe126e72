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


# ------------------------------------------------------------------ GPU box: reference Python on native modules
REF_ZIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "dpvo_ref_py.zip")
_NATIVE = ("cuda_corr", "cuda_ba", "lietorch_backends")


def staged():
    return os.path.exists(REF_ZIP)


def read_config(name):
    """the reference's config/<name>.yaml as a dict (from the staged zip)"""
    import zipfile
    import yaml
    with zipfile.ZipFile(REF_ZIP) as z:
        return yaml.safe_load(z.read("config/%s.yaml" % name))


@contextlib.contextmanager
def reference_python(native=None):
    """Import the reference's UNMODIFIED Python package `dpvo` (packed by oracle/build_ref.py:stage_python into
    oracle/_ref/dpvo_ref_py.zip) with the three native module names bound to `native` =
    (cuda_corr, cuda_ba, lietorch_backends) -- by default OURS (dpvo_b200/_ext), i.e. the drop-in that
    INTEGRATION.md describes, exercised for real.  Third-party packages missing from the image
    (torch_scatter, pypose, matplotlib, yacs) come from oracle/shims.  Needs a CUDA device
    (dpvo/dpvo.py:17 allocates on "cuda" at import).  sys.modules / sys.path are restored on exit."""
    if not staged():
        raise RuntimeError("oracle/_ref/dpvo_ref_py.zip missing: run python oracle/build_ref.py where /root/reference is mounted")
    import dpvo_b200
    if native is None:
        native = dpvo_b200.extensions()[:3]
    saved_path = list(sys.path)
    names = _NAMES + ("lietorch", "gradcheck", "run_tests")
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in names}
    for k in list(saved_mods):
        del sys.modules[k]
    for nm, mod in zip(_NATIVE, native):
        sys.modules[nm] = mod
    sys.path[:0] = [REF_ZIP, SHIMS]
    try:
        yield
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in names]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path


def bind_native(dpvo_pkg_modules, cuda_corr, cuda_ba):
    """re-point an already imported reference package at another pair of native modules (ours <-> oracle/_ref):
    dpvo/altcorr/correlation.py and dpvo/fastba/ba.py look `cuda_corr` / `cuda_ba` up as module globals at call
    time; `neighbors` / `reproject` are bound at import (fastba/ba.py:4-5) and re-exported by fastba/__init__.py"""
    corr_mod = dpvo_pkg_modules["dpvo.altcorr.correlation"]
    ba_mod = dpvo_pkg_modules["dpvo.fastba.ba"]
    fb_pkg = dpvo_pkg_modules["dpvo.fastba"]
    corr_mod.cuda_corr = cuda_corr
    ba_mod.cuda_ba = cuda_ba
    for m in (ba_mod, fb_pkg):
        m.neighbors = cuda_ba.neighbors
        m.reproject = cuda_ba.reproject
