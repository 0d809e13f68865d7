"""CPU oracle for the quantized Conv2d/Linear hot path of LongAoTianxia/Quantized_ViT.

TEST INFRASTRUCTURE ONLY.  Nothing under ``quantized_vit_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do,
and only as the checker / reported baseline - never as the thing shipped.

The oracle restates the reference's algorithm with stock PyTorch-CPU fp32 ops (the reference *is*
stock PyTorch fp32 - SURVEY.md section 0.1 - so the same ATen kernels give the same bits) and NumPy for the
reference's NumPy export path.  Every function cites the reference file:line it follows
(paths relative to the reference root; ``QL`` = QViT_with_GETA/only_train_once/quantization/quant_layers.py,
``QM`` = .../quant_model.py, ``QU`` = "4-bit quantization/quant_ultra.py", ``QZ`` = "4-bit quantization/quantization.py",
``MM`` = "4-bit quantization/mymodel.py", ``VIT`` = QViT_with_GETA/vit_model.py).

PARITY PINNING: the reference ships no golden vectors for low-bit operation.  The oracle
is pinned by (1) the de-facto known-answer vectors of SURVEY.md section 4 (tests/test_oracle_kat.py),
(2) fixtures under ``tests/golden/`` produced by IMPORTING AND RUNNING the unmodified reference in the
build container (``oracle/make_golden.py``, committed), and (3) - when /root/reference is present -
a live cross-check of every oracle function against the reference module it restates
(tests/test_oracle_vs_reference.py).  The reference tree does not exist on the GPU box; nothing in
the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` reads it.
"""
from . import ref_geta, ref_ultra, ref_models  # noqa: F401
