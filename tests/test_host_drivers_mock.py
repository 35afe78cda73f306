"""The HOST DRIVERS of the device paths, executed where no GPU exists: the product's own host code (C ABI, host data model and
csrc/cuda/device_tree.cu: residency, slot arithmetic, pair lists, the refinement loops of add / multiply / refine_grid, dot,
transforms by level) is compiled with g++ against a host-memory stand-in for the CUDA runtime (tests/cpp/cuda_mock), the
kernels are replaced by host functions, fresh "device" memory is poisoned. The GPU tests of the tree algebra and the C++
programs then run against that library. What this cannot check is the kernels themselves -- that is what the -m gpu suite on the
B200 is for -- but every line of driver logic those tests reach is executed here. TEST INFRASTRUCTURE: the mock library is never
shipped and never on the product path."""
import os
import subprocess
import sys

import pytest

import cpp_build as cb


@pytest.fixture(scope="module")
def mock_lib(libs):
    return cb.build_mock_lib()


def test_tree_algebra_gpu_tests_on_the_mock(mock_lib):
    env = dict(os.environ, MRX_TEST_MOCK_LIB=mock_lib)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(cb.ROOT, "tests", "test_zz2_gpu_tree_algebra.py"), "-m", "gpu", "-x", "-q",
                        "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, cwd=cb.ROOT, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout and "skipped" not in r.stdout


def test_projection_and_transform_gpu_tests_on_the_mock(mock_lib):
    """the device projection driver (csrc/cuda/project.cu: refinement loop, pre-built grids), whole-tree transforms, upload /
    download, dot: the cases of the GPU suite that do not depend on the apply kernels' own counters"""
    env = dict(os.environ, MRX_TEST_MOCK_LIB=mock_lib)
    files = [os.path.join(cb.ROOT, "tests", f) for f in ("test_gpu_parity.py", "test_zz1_gpu_reference.py")]
    r = subprocess.run([sys.executable, "-m", "pytest"] + files + ["-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", "-k",
                        "device_projection or identity or golden or vs_real_reference or bottom_up or top_down or hydrogen"],
                       capture_output=True, text=True, env=env, cwd=cb.ROOT, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_node_transform_and_prec_tree_gpu_tests_on_the_mock(mock_lib):
    """the drivers of the standalone node transforms (mrx_node_mw_transform / mrx_node_cv_transform) run for real; the apply with
    precision trees reaches the C ABI and the argument marshalling (its kernels are the B200's business)"""
    env = dict(os.environ, MRX_TEST_MOCK_LIB=mock_lib)
    files = [os.path.join(cb.ROOT, "tests", f) for f in ("test_gpu_node_transforms.py", "test_gpu_prec_trees.py", "test_gpu_periodic.py")]
    r = subprocess.run([sys.executable, "-m", "pytest"] + files + ["-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", "-k", "not needs_a_reach"],
                       capture_output=True, text=True, env=env, cwd=cb.ROOT, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


@pytest.mark.parametrize("program,args", [("apply_drop_in.cpp", ["-1", "all"]), ("scf_hydrogen.cpp", [])])
def test_cpp_programs_on_the_mock(mock_lib, tmp_path, program, args):
    """the C++ programs of the GPU suite linked with the mock library: C++ mirror -> C ABI -> real host drivers -> host kernels"""
    from test_cpp_mirror import check_drop_in_values, check_scf_values
    exe = str(tmp_path / "prog")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(cb.ROOT, "include"), os.path.join(cb.ROOT, "tests", "cpp", program),
           "-o", exe, "-L" + cb.MOCK_DIR, "-lmrcpp_b200_mock", "-Wl,-rpath," + cb.MOCK_DIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = cb.run_program(exe, args=args, env={"MRCPP_B200_DEVICE": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    kv = cb.key_values(r.stdout)
    if program.startswith("scf"):
        check_scf_values(kv)
    else:
        check_drop_in_values(kv, "all")
