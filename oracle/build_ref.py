#!/usr/bin/env python
"""Compile the UNMODIFIED reference CUDA extension into oracle/_ref/.  TEST INFRASTRUCTURE ONLY.

Sources are compiled where they lie under /root/reference/e2edet/module/ops/src (nothing is
copied); the only addition is the force-included oracle/ref_shim.h (one ATen overload the reference
relies on that newer torch dropped).  Output: oracle/_ref/e2edet_ref_ops*.so -- a pybind module with
the reference's own four functions (vision.cpp:7-12), compiled for sm_100a.  It is used

* by tests/test_gpu_vs_reference_cuda.py: our kernels vs the reference's kernels on the same GPU;
* by bench.py as a same-box comparison point (never as the thing measured for `value`).

Runs only where /root/reference exists (the authoring container); the GPU box uses the prebuilt
file, which travels with the repo snapshot (git-ignored, not gpurun-ignored).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/e2edet/module/ops/src"
OUT = os.path.join(HERE, "_ref")
NAME = "e2edet_ref_ops"


def main():
    if not os.path.isdir(SRC):
        print("reference sources not present; nothing to build")
        return 0
    import glob

    if glob.glob(os.path.join(OUT, NAME + "*.so")) and "--force" not in sys.argv:
        print("oracle/_ref already built")
        return 0
    os.makedirs(os.path.join(OUT, "build"), exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    shim = os.path.join(HERE, "ref_shim.h")
    sources = [os.path.join(SRC, "vision.cpp"),
               os.path.join(SRC, "box_attn", "box_attn.cu"),
               os.path.join(SRC, "instance_attn", "instance_attn.cu")]
    load(
        name=NAME,
        sources=sources,
        extra_include_paths=[SRC],
        extra_cflags=["-O2", "-DWITH_CUDA", "-include", shim],   # WITH_CUDA: setup.py:47
        extra_cuda_cflags=["-O3", "-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                           "-D__CUDA_NO_HALF2_OPERATORS__", "-include", shim,
                           "-gencode", "arch=compute_100a,code=sm_100a"],
        build_directory=os.path.join(OUT, "build"),
        is_python_module=False,
        verbose="-v" in sys.argv,
    )
    import shutil

    built = os.path.join(OUT, "build", NAME + ".so")
    shutil.copy2(built, os.path.join(OUT, NAME + ".so"))
    print("built", os.path.join(OUT, NAME + ".so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
