"""Builds librjb200.so (CUDA kernels + C ABI + host CDB code) and the two
command-line tools for sm_100a with nvcc.  In-tree outputs, no JIT cache."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "librjb200.so")
NVCC = os.environ.get("RJB_NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "--expt-extended-lambda",
          "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(HERE, "csrc")]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources():
    csrc = os.path.join(HERE, "csrc")
    host = os.path.join(HERE, "host")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)]
    deps += [os.path.join(host, f) for f in os.listdir(host)]
    deps.append(os.path.join(ROOT, "include", "rjb200.h"))
    return deps


def build(force=False, verbose=False):
    deps = _sources()
    if force or _newer(LIB, deps):
        cmd = [NVCC] + ARCH + COMMON + os.environ.get("RJB_DEFINES", "").split() + \
              ["-Xcompiler", "-fPIC,-pthread", "-shared", "-o", os.environ.get("RJB_LIB_OUT", LIB),
                                        os.path.join(HERE, "csrc", "rjb_api.cu"),
                                        os.path.join(HERE, "host", "cdb.cc")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    bindir = os.path.join(HERE, "bin")
    os.makedirs(bindir, exist_ok=True)
    for exe, src in (("query_exec", "query_exec.cc"), ("polyover_exec", "polyover_exec.cc")):
        srcp = os.path.join(HERE, "host", src)
        out = os.path.join(bindir, exe)
        if not os.path.exists(srcp):
            continue
        if force or _newer(out, deps + [LIB]):
            cmd = [NVCC] + COMMON + ["-o", out, srcp, "-L" + HERE, "-lrjb200",
                                     "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/.."]
            subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
