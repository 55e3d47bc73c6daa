"""g++ build of the host library (`libstep50_b200.so`) and the `main` executable (both link the CUDA library) and of
`libministep_b200.so`: ministep + its C shim alone, no CUDA dependency (mesh / numbering / host assembly for the CPU
arm of bench.py and the CPU tests)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "lib")
CXX = os.environ.get("GMG_CXX", "/usr/bin/g++")  # the image's $CXX (/opt/gcc) has no libgomp
FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fopenmp", "-Wall", "-Wno-unknown-pragmas"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False):
    os.makedirs(LIB, exist_ok=True)
    srcs = [os.path.join(HERE, f) for f in ("ministep.cc", "step-50.cc", "capi_ministep.cc", "capi_host.cc")]
    ms_srcs = [os.path.join(HERE, f) for f in ("ministep.cc", "capi_ministep.cc")]
    ms_deps = ms_srcs + [os.path.join(HERE, f) for f in ("ministep.h", "capi_ministep.h")] + [
        os.path.join(PKG, "csrc", "assemble_row.h")]
    ms_lib = os.path.join(LIB, "libministep_b200.so")
    if force or _stale(ms_lib, ms_deps):
        cmd = [CXX] + FLAGS + ["-shared", "-o", ms_lib] + ms_srcs
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cc"))]
    deps.append(os.path.join(LIB, "libgmg_b200.so"))
    lib = os.path.join(LIB, "libstep50_b200.so")
    exe = os.path.join(LIB, "main")
    link = ["-L" + LIB, "-lgmg_b200", "-Wl,-rpath,$ORIGIN"]
    if force or _stale(lib, deps):
        cmd = [CXX] + FLAGS + ["-shared", "-o", lib] + srcs + link
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    if force or _stale(exe, deps + [lib]):
        cmd = [CXX] + FLAGS + ["-o", exe, os.path.join(HERE, "main.cc"), "-L" + LIB, "-lstep50_b200", "-lgmg_b200",
                               "-Wl,-rpath,$ORIGIN"]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return [lib, exe, ms_lib]


if __name__ == "__main__":
    print(build(force=True))
