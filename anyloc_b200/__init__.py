"""anyloc_b200 -- B200-native (sm_100a) implementation of AnyLoc's DINOv2 -> VLAD -> top-k hot path.

`anyloc_b200.utilities` mirrors the reference's `utilities.py` API for that path; the arithmetic
lives in `libanyloc_b200.so` (C ABI: include/anyloc_b200.h).  Importing this package does not need
a GPU; every compute call does (there is no CPU fallback)."""
__version__ = "0.1.0"
