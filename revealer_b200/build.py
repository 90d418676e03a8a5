"""In-tree build of librevealer_b200.so (sm_100a only) with nvcc.  No JIT cache: the .so sits next
to this file so that it travels to the GPU box with the repo snapshot."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "librevealer_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]
# (source, extra flags): bandwidth.cu must not contract a*b+c into FMA (see its header)
UNITS = [
    ("bandwidth.cu", ["-fmad=false"]),
    ("score.cu", []),
    ("gct.cu", []),
    ("hostpack.cpp", ["-Xcompiler", "-O3"]),   # plain host C++ (g++ through nvcc): packer threads of the cooperative upload
    ("capi.cu", []),
]
DEPS = ["rvl_common.cuh", os.path.join("..", "..", "include", "revealer_b200.h")]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build librevealer_b200.so")
    return nvcc


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    env = dict(os.environ)
    # $CC/$CXX in this image point at a wrapper; let nvcc use the system g++
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objs = []
    deps = [os.path.join(CSRC, d) for d in DEPS] + [os.path.abspath(__file__)]
    rebuilt = False
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        if force or _stale(o, [s] + deps):
            cmd = [nvcc] + ccbin + ARCH + COMMON + extra + ["-Xptxas", "-v", "-c", s, "-o", o]
            r = subprocess.run(cmd, env=env, capture_output=True, text=True)
            if verbose or r.returncode:
                print(" ".join(cmd))
                print(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed for " + src)
            with open(o + ".ptxas.txt", "w") as fh:
                fh.write(r.stderr)
            rebuilt = True
        objs.append(o)
    if rebuilt or force or _stale(SO, objs):
        cmd = [nvcc] + ccbin + ARCH + ["-shared", "-o", SO] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if verbose or r.returncode:
            print(" ".join(cmd))
            print(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("link failed")
    return SO


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose=True))
