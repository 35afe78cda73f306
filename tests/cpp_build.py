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
