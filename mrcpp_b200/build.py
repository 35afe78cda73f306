"""In-tree build of libmrcpp_b200.so (CUDA kernels + C-ABI + host data model) for sm_100a, and of
the CPU oracle (oracle/_build/liboracle.so). nvcc cross-compiles without a GPU."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "lib")
LIB = os.path.join(LIBDIR, "libmrcpp_b200.so")
ORACLE_DIR = os.path.join(os.path.dirname(ROOT), "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

SOURCES = [
    "cabi.cpp",
    "host/tables.cpp",
    "host/tree.cpp",
    "host/operators.cpp",
    "cuda/device_tree.cu",
    "cuda/kernels.cu",
    "cuda/project.cu",
    "cuda/apply.cu",
    "cuda/apply_kernels.cu",
    "cuda/apply_pipeline.cu",
    "cuda/apply_enum.cu",
    "cuda/apply_split.cu",
    "cuda/apply_prec.cu",
    "cuda/comm.cu",
    "cuda/microbench.cu",
]
HEADERS = ["engine.hpp", "host/mrx_host.hpp", "cuda/common.cuh", "cuda/kernels.cuh", "cuda/apply_kernels.cuh",
           "../../include/mrcpp_b200.h"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-ccbin", "/usr/bin/g++",
    "-Xcompiler", "-fPIC,-fopenmp,-O3,-march=x86-64-v3",
    "-Xptxas", "-v",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace("/", "_") + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + ["-x", "cu", "-dc" if False else "-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(f"--- {s}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
        with open(os.path.join(objdir, s.replace("/", "_") + ".log"), "w") as f:
            f.write(out)
    if force or procs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-Xcompiler", "-fopenmp", "-lgomp", "-ldl", "-cudart", "static"]
        subprocess.check_call(cmd, stdout=sys.stderr)
    return LIB


def build_oracle(force=False):
    # make's chatter goes to stderr: bench.py must print exactly one JSON line on stdout
    if force:
        subprocess.call(["make", "-s", "-C", ORACLE_DIR, "clean"], stdout=sys.stderr)
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=sys.stderr)
    return ORACLE_LIB


def build_ref(reference="/root/reference"):
    """oracle/_ref: the reference's own sources compiled in place (oracle/build_ref.sh) -- only where the reference tree
    exists (this container); the GPU box uses the prebuilt files that travel with the snapshot. Returns the driver library
    path or None."""
    out = os.path.join(ORACLE_DIR, "_ref", "libref_driver.so")
    if os.path.isdir(os.path.join(reference, "src")):
        subprocess.check_call(["bash", os.path.join(ORACLE_DIR, "build_ref.sh"), reference], stdout=sys.stderr)
    return out if os.path.exists(out) else None


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_lib(force=force, verbose="-v" in sys.argv))
    print(build_oracle(force=force))
    print(build_ref())
