"""Compile the reference's own ``sparse_conv_ext`` from the sources WHERE THEY LIE under
/root/reference into the git-ignored ``oracle/_ref/``.  TEST INFRASTRUCTURE ONLY.

Nothing is copied into the repo: only the built shared object (and ninja's object files) land in
``oracle/_ref/``; it travels to the GPU box with the gpurun snapshot.  The recipe is the one of
/root/reference/setup.py:89-110 with a single flag changed: ``-std=c++17`` instead of ``-std=c++14``
(torch >= 2.1 headers refuse C++14).  The reference's own build system is not run.

Usage:  python oracle/build_ref.py            (no-op when /root/reference is absent)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("FV2P_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
EXT_NAME = "sparse_conv_ext"


def built_path():
    if not os.path.isdir(OUT_DIR):
        return None
    for f in sorted(os.listdir(OUT_DIR)):
        if f.startswith(EXT_NAME) and f.endswith(".so"):
            return os.path.join(OUT_DIR, f)
    return None


def build(verbose=False):
    src_root = os.path.join(REF_ROOT, "pcdet", "ops", "spconv")
    if not os.path.isdir(src_root):
        return built_path()
    if built_path() is not None:
        return built_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load

    files = ["all.cc", "reordering.cc", "reordering_cuda.cu", "indice.cc", "indice_cuda.cu", "maxpool.cc",
             "maxpool_cuda.cu"]  # setup.py:100-108
    defs = ["-w", "-std=c++17", "-DWITH_CUDA"]
    load(name=EXT_NAME, sources=[os.path.join(src_root, "src", f) for f in files],
         extra_include_paths=[os.path.join(src_root, "include")], extra_cflags=defs,
         extra_cuda_cflags=defs + ["-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                                   "-D__CUDA_NO_HALF2_OPERATORS__"],  # setup.py:43-47
         with_cuda=True, build_directory=OUT_DIR, verbose=verbose, is_python_module=False)
    return built_path()


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("reference extension:", p if p else "unavailable (no /root/reference and nothing prebuilt)")
