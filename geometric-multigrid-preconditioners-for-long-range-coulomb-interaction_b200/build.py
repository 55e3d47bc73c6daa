"""Build the in-tree native libraries with nvcc / g++ (no JIT cache, the .so files travel with the tree).

    python build.py            # libgmg_b200.so (CUDA, sm_100a) + host library + main
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler",
              "-fPIC,-O3", "-shared", "-cudart", "static"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(verbose=False, force=False):
    os.makedirs(LIB, exist_ok=True)
    src_dir = os.path.join(HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in ("context.cu", "rhs.cu", "indicator.cu", "partition.cc")]
    deps = [os.path.join(src_dir, f) for f in os.listdir(src_dir)] + [os.path.join(ROOT, "include", "gmg_b200.h")]
    out = os.path.join(LIB, "libgmg_b200.so")
    if force or _stale(out, deps):
        cmd = [NVCC] + CUDA_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + srcs
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return out


def build_all(verbose=False, force=False):
    outs = [build_cuda(verbose, force)]
    host = os.path.join(HERE, "host", "build_host.py")
    if os.path.exists(host):
        import importlib.util
        spec = importlib.util.spec_from_file_location("build_host", host)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        outs += mod.build(force=force)
    return outs


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv, force="-f" in sys.argv))
