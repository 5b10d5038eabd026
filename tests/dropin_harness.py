"""TEST INFRASTRUCTURE -- runs the reference's UNMODIFIED driver, `build_vlads` of
/root/reference/scripts/dino_v2_vlad.py:124-303 (extract loop :164-188, vocabulary :195-213, database / query
VLADs :219-264), against a synthetic dataset object, with a chosen module answering `from utilities import ...`:
either the reference's own utilities.py (verbatim, behind the oracle's stand-ins for faiss / fpk / the hub) or this
repo's drop-in shim (anyloc_b200/dropin/utilities.py).  Nothing of the reference is copied: the script is imported
from where it lies, so this only works where /root/reference exists (the build container).  Third-party modules the
script imports but never uses on this path (natsort, matplotlib, faiss in the dataset loaders) are stubbed.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

from oracle import reference_import as ri

SCRIPT = os.path.join(ri.REFERENCE_ROOT, "scripts", "dino_v2_vlad.py")


def available():
    return os.path.isfile(SCRIPT)


class SyntheticVprDataset:
    """What build_vlads needs from a BaseDataset (dvgl_benchmark/datasets_ws.py:222-239): `database_num`, `len()`,
    `ds[i][0]` = normalised image tensor [3,h,w], `get_image_relpaths(indices)`, `soft_positives_per_query`."""

    def __init__(self, n_db=6, n_qu=3, h=60, w=75, seed=5):
        g = torch.Generator().manual_seed(seed)
        self.database_num, self.queries_num = n_db, n_qu
        db = torch.randn(n_db, 3, h, w, generator=g)
        qu = db[:n_qu] + 0.05 * torch.randn(n_qu, 3, h, w, generator=g)       # query i shows database place i
        self.images = torch.cat([db, qu])
        self.soft_positives_per_query = np.empty(n_qu, dtype=object)
        for i in range(n_qu):
            self.soft_positives_per_query[i] = np.array([i])

    def __len__(self):
        return self.images.shape[0]

    def __getitem__(self, i):
        return self.images[i], i

    def get_image_relpaths(self, i):
        if isinstance(i, (int, np.integer)):
            return f"synth/img_{int(i):04d}.jpg"
        return [f"synth/img_{int(k):04d}.jpg" for k in i]


def load_script(utilities_module):
    """Imports scripts/dino_v2_vlad.py (unmodified) with `utilities` resolving to `utilities_module`."""
    ri._install_stubs()
    if "natsort" not in sys.modules:
        ns = types.ModuleType("natsort")
        ns.natsorted = sorted
        sys.modules["natsort"] = ns
    if ri.REFERENCE_ROOT not in sys.path:
        sys.path.append(ri.REFERENCE_ROOT)
    old = sys.modules.get("utilities")
    sys.modules["utilities"] = utilities_module
    # the dataset loaders do `from utilities import CustomDataset` at import time: drop cached copies bound to another module
    for name in [m for m in sys.modules if m.startswith("custom_datasets") or m.startswith("dvgl_benchmark")]:
        del sys.modules[name]
    try:
        spec = importlib.util.spec_from_file_location("_ref_dino_v2_vlad_" + utilities_module.__name__.replace(".", "_"), SCRIPT)
        mod = importlib.util.module_from_spec(spec)
        st = (np.random.get_state(), torch.random.get_rng_state())
        spec.loader.exec_module(mod)
        np.random.set_state(st[0]); torch.random.set_rng_state(st[1])
    finally:
        if old is not None:
            sys.modules["utilities"] = old
        else:
            sys.modules.pop("utilities", None)
    return mod


def make_largs(mod, cache_dir, model="dinov2_vits14", layer=2, facet="value", clusters=4, cache=False, soft=False):
    prog = type(mod.LocalArgs().prog)(cache_dir=cache_dir, vg_dataset_name="17places", use_wandb=False)
    return mod.LocalArgs(prog=prog, model_type=model, desc_layer=layer, desc_facet=facet, num_clusters=clusters,
                         cache_vlad_descs=cache, vlad_assignment="soft" if soft else "hard")
