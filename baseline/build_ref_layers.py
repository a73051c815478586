#!/usr/bin/env python
"""Compile the UNMODIFIED reference transformer layers to bytecode under baseline/_ref/.

The drop-in claim is that BoxeR's own layers -- ``BoxTransformer`` (e2edet/module/box_transformer.py) and
``Box3dTransformer`` (e2edet/module/box3d_transformer.py) -- run unchanged on ``boxer_b200``.  Testing and timing
that on the GPU box needs those layers there, and /root/reference exists only in the authoring container.  Nothing
from the reference is copied into the repo: the files below are compiled where they lie (``py_compile``), the
outputs are sourceless ``.pyc`` files under baseline/_ref/ (git-ignored; they travel to the GPU box with the
snapshot like the built ``.so`` files), laid out as the namespace packages ``e2edet.module`` / ``e2edet.utils`` --
without the reference's package ``__init__`` files, which pull in the trainer, datasets, omegaconf and pycocotools.

    box_transformer.py, box3d_transformer.py     the callers under test
    utils/general.py (+ distributed.py, box_ops.py it imports)   their helpers (flatten_with_shape, get_clones, ...)
    box_attention.py, ops/box_attention_func.py  the reference's OWN attention modules, used only as the comparison
                                                 side of tests/test_gpu_reference_layers.py (backed by the CPU oracle)

Runs only where /root/reference exists; python -m baseline.build_ref_layers [--force]
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/e2edet"
OUT = os.path.join(HERE, "_ref", "e2edet")
FILES = [
    "module/box_transformer.py",
    "module/box3d_transformer.py",
    "module/box_attention.py",
    "module/ops/box_attention_func.py",
    "utils/general.py",
    "utils/distributed.py",
    "utils/box_ops.py",
]


def main() -> int:
    if not os.path.isdir(REF):
        print("reference sources not present; nothing to build")
        return 0
    force = "--force" in sys.argv
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, rel[:-3] + ".pyc")
        if os.path.exists(dst) and not force and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile="e2edet/" + rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        print("compiled", rel)
    return 0


if __name__ == "__main__":
    sys.exit(main())
