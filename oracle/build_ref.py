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


PN2_NAME = "libref_pointnet2.so"


def pointnet2_path():
    p = os.path.join(OUT_DIR, PN2_NAME)
    return p if os.path.exists(p) else None


def build_pointnet2(verbose=False):
    """The reference's three_nn / three_interpolate (pointnet2_batch/src/interpolate_gpu.cu:16-139) and voxel_query
    (pointnet2_stack/src/voxel_query_gpu.cu:10-121) CUDA kernels with their host launchers, compiled with nvcc from
    the .cu files where they lie into oracle/_ref/libref_pointnet2.so.  Their pybind wrappers (*.cpp) include
    THC/THC.h, which torch >= 2 no longer ships, so the torch extension itself is unbuildable here; the launchers are
    plain C++ functions over device pointers and are called through ctypes by their mangled names (oracle/ref.py)."""
    root = os.path.join(REF_ROOT, "pcdet", "ops", "pointnet2")
    if not os.path.isdir(root):
        return pointnet2_path()
    if pointnet2_path() is not None:
        return pointnet2_path()
    import subprocess
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # the kernels' own headers include <torch/serialize/tensor.h> for the (unused here) wrapper declarations
    import sysconfig
    from torch.utils.cpp_extension import include_paths
    incs = []
    for d in include_paths() + [sysconfig.get_paths()["include"]]:
        incs += ["-I", d]
    objs = []
    for sub, name in (("pointnet2_batch", "interpolate_gpu"), ("pointnet2_stack", "voxel_query_gpu")):
        src = os.path.join(root, sub, "src", name + ".cu")
        obj = os.path.join(OUT_DIR, "pn2_%s_%s.o" % (sub, name))
        cmd = [nvcc, "-gencode", "arch=compute_100,code=sm_100", "-O3", "-w", "-std=c++17", "-Xcompiler", "-fPIC",
               "-I", os.path.join(root, sub, "src")] + incs + ["-c", src, "-o", obj]
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or out.returncode:
            sys.stderr.write(out.stdout)
        if out.returncode:
            return None
        objs.append(obj)
    out = subprocess.run([nvcc, "-gencode", "arch=compute_100,code=sm_100", "-shared", "-o",
                          os.path.join(OUT_DIR, PN2_NAME)] + objs + ["-Xcompiler", "-fPIC"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode:
        sys.stderr.write(out.stdout)
        return None
    return pointnet2_path()


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("reference extension:", p if p else "unavailable (no /root/reference and nothing prebuilt)")
    q = build_pointnet2(verbose="-v" in sys.argv)
    print("reference pointnet2 kernels:", q if q else "unavailable")
