"""`utilities` shim: put this directory FIRST on PYTHONPATH (before the AnyLoc checkout) and the
reference's scripts (`from utilities import VLAD, get_top_k_recall, DinoV2ExtractFeatures, ...`,
scripts/dino_v2_vlad.py:37-38,50) run on anyloc_b200 unchanged."""
import os as _os
import sys as _sys

_repo = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _repo not in _sys.path:
    _sys.path.insert(0, _repo)

from anyloc_b200.utilities import *  # noqa: F401,F403,E402
from anyloc_b200.utilities import (  # noqa: F401,E402
    VLAD, DinoV2ExtractFeatures, get_top_k_recall, seed_everything, reduce_pca, CustomDataset, to_np,
    top_k_search, to_pil_list, pad_img, concat_desc_dists_clusters, FlatIndex, _DINO_V2_MODELS, _DINO_FACETS)
