"""Builds libliodom_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libliodom_b200.so")

# translation unit -> extra flags.  The bit-exact stages forbid FMA contraction.
UNITS = {
    "extract.cu": ["-fmad=false"],
    "register.cu": ["-fmad=false"],
    "voxelgrid.cu": ["-fmad=false"],
    "solve.cu": [],
    "map.cu": ["-fmad=false"],
    "shard.cu": [],
    "cabi.cu": [],
}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", "g++"] + os.environ.get("LIODOM_NVCC_EXTRA", "").split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "liodom_b200.h"))
    objs = []
    procs = []
    for cu, extra in UNITS.items():
        src = os.path.join(CSRC, cu)
        if not os.path.exists(src):
            continue
        obj = os.path.join(CSRC, cu[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = ["nvcc"] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((cu, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cu, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (cu, out))
        elif verbose or "warning" in out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("CUDA build failed")
    if force or _stale(SO, objs):
        subprocess.check_call(["nvcc"] + ARCH + ["-shared", "-ccbin", "g++", "-o", SO] + objs + ["-ldl"])
    build_host(force)
    return SO


HOST_SO = os.path.join(HERE, "libliodom_host.so")


def build_host(force=False):
    """The C++ facade (reference class surface over the C ABI); links against the CUDA library."""
    inc = os.path.join(HERE, "..", "include")
    src = os.path.join(HERE, "host", "facade.cc")
    deps = [src, SO] + [os.path.join(inc, "liodom", f) for f in os.listdir(os.path.join(inc, "liodom"))]
    if force or _stale(HOST_SO, deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-I", inc, "-o", HOST_SO, src,
                               "-L", HERE, "-lliodom_b200", "-Wl,-rpath,$ORIGIN", "-lpthread"])
    return HOST_SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
