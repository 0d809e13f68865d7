"""Copy the reference's own hot-path files (pure Python, nothing to compile) into oracle/_ref/ so that the UNMODIFIED
reference implementation can be timed on the GPU box's host cores by `bench.py --impl reference` / `cpu_baseline`
(kind "reference").  oracle/_ref/ is git-ignored (never part of the history) but travels with the gpurun snapshot, like
the built .so files.  Runs in the build container only (needs /root/reference); __graft_entry__.build() calls it.

    python oracle/make_ref.py
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("QVIT_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = [
    "QViT_with_GETA/vit_model.py",
    "QViT_with_GETA/only_train_once/quantization/__init__.py",
    "QViT_with_GETA/only_train_once/quantization/quant_layers.py",
    "QViT_with_GETA/only_train_once/quantization/quant_model.py",
    "4-bit quantization/quant_ultra.py",
    "4-bit quantization/quantization.py",
    "4-bit quantization/mymodel.py",
]


def make_ref() -> bool:
    if not os.path.isdir(SRC):
        return False
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    return True


if __name__ == "__main__":
    ok = make_ref()
    print("oracle/_ref written" if ok else f"{SRC} not found: nothing copied")
    sys.exit(0 if ok else 1)
