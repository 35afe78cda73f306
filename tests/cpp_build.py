"""Compile and run the C++ programs under tests/cpp/ against the C++ MRCPP mirror (include/MRCPP/) and libmrcpp_b200.so."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mrcpp_b200", "lib")
TABLES = os.path.join(ROOT, "mrcpp_b200", "data", "mwtables.bin")
ORACLE = os.path.join(ROOT, "oracle", "_build", "liboracle.so")


def compile_program(sources, out, extra=(), werror=True):
    """werror=False for the reference's own example sources (they carry warnings of their own)"""
    cmd = ["g++", "-std=c++17", "-O2", "-Wall"] + (["-Werror"] if werror else []) + ["-rdynamic", "-I" + os.path.join(ROOT, "include")] + list(sources) + \
          ["-o", out, "-L" + LIBDIR, "-lmrcpp_b200", "-ldl", "-Wl,-rpath," + LIBDIR] + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def run_program(exe, args=(), env=None, timeout=600):
    e = dict(os.environ)
    e["MRX_TABLES"] = TABLES
    e["MRX_TEST_ORACLE"] = ORACLE
    e.update(env or {})
    return subprocess.run([exe] + list(args), capture_output=True, text=True, env=e, timeout=timeout, cwd=ROOT)


def key_values(stdout):
    """'key value' lines of a test program -> dict of floats"""
    out = {}
    for line in stdout.splitlines():
        parts = line.split()
        if len(parts) == 2:
            try:
                out[parts[0]] = float(parts[1])
            except ValueError:
                pass
    return out


MOCK_DIR = os.path.join(ROOT, "tests", "_mock")
MOCK_LIB = os.path.join(MOCK_DIR, "libmrcpp_b200_mock.so")


def build_mock_lib():
    """libmrcpp_b200_mock.so: the product's HOST code (C ABI, host data model, the host drivers of csrc/cuda/device_tree.cu)
    compiled with g++ against the host-memory CUDA stand-in of tests/cpp/cuda_mock, kernels replaced by host functions. TEST
    INFRASTRUCTURE: lets the driver logic of the tree algebra run where no GPU exists."""
    csrc = os.path.join(ROOT, "mrcpp_b200", "csrc")
    mock = os.path.join(ROOT, "tests", "cpp", "cuda_mock")
    srcs = [os.path.join(csrc, "cabi.cpp"), os.path.join(csrc, "host", "tables.cpp"), os.path.join(csrc, "host", "tree.cpp"),
            os.path.join(csrc, "host", "operators.cpp"), os.path.join(csrc, "cuda", "device_tree.cu"), os.path.join(csrc, "cuda", "project.cu"), os.path.join(mock, "mock_kernels.cpp")]
    deps = srcs + [os.path.join(csrc, "cuda", "kernels.cu"), os.path.join(mock, "cuda_runtime.h"), os.path.join(csrc, "engine.hpp"), os.path.join(csrc, "host", "mrx_host.hpp"),
                   os.path.join(ROOT, "oracle", "oracle.cpp"), os.path.join(ROOT, "include", "mrcpp_b200.h")]
    if os.path.exists(MOCK_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(MOCK_LIB) for d in deps):
        return MOCK_LIB
    os.makedirs(MOCK_DIR, exist_ok=True)
    # the element-wise kernels of the tree algebra have no cross-thread communication: their SOURCE is cut out of kernels.cu and
    # compiled for the host, where mock_kernels.cpp runs it block by block, thread by thread (blockIdx / threadIdx as variables)
    ksrc = open(os.path.join(csrc, "cuda", "kernels.cu")).read()
    with open(os.path.join(MOCK_DIR, "extracted_kernels.inc"), "w") as f:
        for name in ("axpy_nodes_kernel", "product_values_kernel"):
            start = ksrc.index("__global__ void __launch_bounds__(256) " + name)
            end = ksrc.index("\n}\n", start) + 3
            f.write(ksrc[start:end] + "\n")
    objs, procs = [], []
    for src in srcs:
        obj = os.path.join(MOCK_DIR, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = ["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-w", "-I" + mock, "-I" + MOCK_DIR, "-x", "c++", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        assert p.returncode == 0, f"{src}:\n{out}"
    r = subprocess.run(["g++", "-shared", "-fopenmp", "-o", MOCK_LIB] + objs + ["-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return MOCK_LIB
