"""TEST INFRASTRUCTURE -- Tier-A oracle: import the reference's own
/root/reference/utilities.py VERBATIM (nothing is copied into this repo) behind
stand-ins for the three third-party packages it imports but that are not
installed here (fast_pytorch_kmeans, faiss, matplotlib).  Only usable in the
build container (where /root/reference exists); used by
tests/golden/make_golden.py to generate the committed golden vectors and by the
`-m "not gpu"` tests to pin oracle/anyloc_oracle.py when the reference is
present.  Nothing on the GPU box may call this.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "utilities.py"))


def _install_stubs():
    from oracle import fpk_restated, faiss_restated
    if "fast_pytorch_kmeans" not in sys.modules:
        m = types.ModuleType("fast_pytorch_kmeans")
        m.KMeans = fpk_restated.KMeans
        sys.modules["fast_pytorch_kmeans"] = m
    if "faiss" not in sys.modules:
        f = types.ModuleType("faiss")
        for n in ("IndexFlatIP", "IndexFlatL2", "StandardGpuResources", "index_cpu_to_gpu"):
            setattr(f, n, getattr(faiss_restated, n))
        c = types.ModuleType("faiss.contrib")
        t = types.ModuleType("faiss.contrib.torch_utils")
        f.contrib = c
        c.torch_utils = t
        sys.modules["faiss"] = f
        sys.modules["faiss.contrib"] = c
        sys.modules["faiss.contrib.torch_utils"] = t
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mp = types.ModuleType("matplotlib")
        pp = types.ModuleType("matplotlib.pyplot")
        mp.pyplot = pp
        sys.modules["matplotlib"] = mp
        sys.modules["matplotlib.pyplot"] = pp


_cached = None


def load_reference_utilities():
    """Returns the module object of /root/reference/utilities.py (imported under
    the private name `_anyloc_reference_utilities` so it never shadows the
    product's own `utilities` shim)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not present (only in the build container)")
    _install_stubs()
    spec = importlib.util.spec_from_file_location(
        "_anyloc_reference_utilities", os.path.join(REFERENCE_ROOT, "utilities.py"))
    mod = importlib.util.module_from_spec(spec)
    import numpy as np, random, torch
    st = (random.getstate(), np.random.get_state(), torch.random.get_rng_state())
    spec.loader.exec_module(mod)          # runs seed_everything() (utilities.py:1011)
    random.setstate(st[0]); np.random.set_state(st[1]); torch.random.set_rng_state(st[2])
    _cached = mod
    return mod


class hub_patched:
    """Context manager: torch.hub.load('facebookresearch/dinov2', name) returns the
    restated model (oracle/dinov2_restated.py), so the reference's unmodified
    DinoV2ExtractFeatures (utilities.py:223-285) runs offline."""

    def __init__(self, model_factory):
        self.factory = model_factory

    def __enter__(self):
        import torch
        self._orig = torch.hub.load
        torch.hub.load = lambda repo, name, *a, **k: self.factory(name)
        return self

    def __exit__(self, *exc):
        import torch
        torch.hub.load = self._orig
