"""oracle/ -- TEST INFRASTRUCTURE ONLY (not product code).

CPU restatement of the AnyLoc DINOv2 -> hard-VLAD -> cosine top-k hot path
(reference: /root/reference/utilities.py:219-288, 390-469, 624-1008).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import anything from this package, and only as the
*checker* or the *timed CPU baseline* -- never on the product path.  The
product (anyloc_b200/) fails loudly when its CUDA library is missing; it has
no CPU fallback and never imports oracle/.

Parity status ("pinning"):
  * VLAD.generate / generate_multi / get_top_k_recall: the restatement in
    oracle/anyloc_oracle.py is PINNED against the reference's own code,
    imported verbatim from /root/reference/utilities.py in the build
    container (oracle/reference_import.py) -- golden vectors committed under
    tests/golden/ together with the generating script
    (tests/golden/make_golden.py).
  * The arithmetic of three third-party packages the reference calls but does
    not vendor is restated from their published behaviour:
      - facebookresearch/dinov2 @ main (unpinned upstream)  -> oracle/dinov2_restated.py
        cross-checked against the independent HuggingFace port
        (transformers.models.dinov2) present in this image;
      - fast-pytorch-kmeans==0.1.6                           -> oracle/fpk_restated.py
      - faiss-gpu==1.7.2 (IndexFlatIP / IndexFlatL2)         -> oracle/faiss_restated.py
    The reference holds NO tests, golden vectors or fixtures for any of these
    (SURVEY.md section 4), so at those three boundaries parity is
    "PARITY UNPINNED" in the sense of the task statement: anchored on the
    reference's call sites + the published algorithms (+ the HF cross-check
    for the ViT), not on reference-held vectors.
"""
