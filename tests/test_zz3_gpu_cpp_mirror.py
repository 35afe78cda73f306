"""GPU test (-m gpu) of the C++ host mirror (include/MRCPP/): tests/cpp/apply_drop_in.cpp -- a program written against the
MRCPP API names -- is compiled with g++, linked with libmrcpp_b200.so and run on the device; its numbers must reproduce the
reference's known answers and the same cases run through the Python mirror (same C ABI, same kernels)."""
import math
import os

import pytest

import cpp_build as cb
from test_cpp_mirror import check_drop_in_values, check_periodic_values, check_scf_values

pytestmark = pytest.mark.gpu


def _run_drop_in(tmp_path, what):
    from mrcpp_b200 import _lib
    if _lib.device() is None or _lib.device() < 0:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    exe = cb.compile_program([os.path.join(cb.ROOT, "tests", "cpp", "apply_drop_in.cpp")], str(tmp_path / "apply_drop_in"))
    r = cb.run_program(exe, args=["-1", what], env={"MRCPP_B200_DEVICE": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    kv = cb.key_values(r.stdout)
    assert kv["done"] == 1
    check_drop_in_values(kv, what)
    return kv


def test_drop_in_core_on_the_device(libs, tmp_path):
    """the apply path through the C++ mirror: Poisson apply, hydrogen fixed point of the Helmholtz operator, ABGV derivative"""
    mw, orc = libs
    kv = _run_drop_in(tmp_path, "core")
    assert kv["poisson_launches"] > 0  # CUDA kernels ran inside the apply
    # same Poisson case through the Python mirror on the device
    k, prec, beta = 7, 1e-5, 100.0
    mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
    f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)
    ft, gt = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, ft, f, device=True)
    st = mw.apply(prec, gt, mw.PoissonOperator(mra, prec), ft)
    assert kv["poisson_f_nodes"] == ft.getNNodes() and kv["poisson_g_nodes"] == gt.getNNodes()
    assert kv["poisson_tuples"] == st.f_applied and kv["poisson_calc_nodes"] == st.g_nodes
    en = mw.dot(gt, ft)
    assert abs(kv["poisson_energy"] - en) <= 1e-12 * abs(en)
    assert abs(kv["poisson_g_sqnorm"] - gt.getSquareNorm()) <= 1e-12 * gt.getSquareNorm()
    assert abs(kv["poisson_f_integral"] - ft.integrate()) <= 1e-13 and abs(kv["poisson_g_integral"] - gt.integrate()) <= 1e-9 * abs(gt.integrate())


def test_drop_in_algebra_on_the_device(libs, tmp_path):
    """the callers around the apply through the C++ mirror: gradient, divergence, add (fixed grid and adaptive), multiply,
    square, evalf"""
    _run_drop_in(tmp_path, "algebra")


def test_scf_cycle_on_the_device(libs, tmp_path):
    """tests/cpp/scf_hydrogen.cpp: a complete self-consistency loop on resident trees (projection, multiply, Helmholtz operator
    construction + apply, rescale, add, dot, normalize) on the device; same energies as the run of the same program with the
    device entry points served by the CPU oracle"""
    src = os.path.join(cb.ROOT, "tests", "cpp", "scf_hydrogen.cpp")
    exe = cb.compile_program([src], str(tmp_path / "scf_gpu"))
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    kv = cb.key_values(r.stdout)
    check_scf_values(kv)
    cpu = cb.compile_program([src, os.path.join(cb.ROOT, "tests", "cpp", "oracle_backend.cpp")], str(tmp_path / "scf_cpu"))
    rc = cb.run_program(cpu, env={"MRCPP_B200_DEVICE": "-1"})
    assert rc.returncode == 0, rc.stderr[-2000:]
    kc = cb.key_values(rc.stdout)
    assert kv["iterations"] == kc["iterations"]
    # the two runs differ in rounding only (same algorithm, same thresholds); a borderline split decision may still fall
    # differently over seven chained operations, so node counts are compared loosely and energies at 1e-6
    for it in range(1, int(kv["iterations"]) + 1):
        assert abs(kv[f"nodes_{it}"] - kc[f"nodes_{it}"]) <= 0.02 * kc[f"nodes_{it}"]
        assert abs(kv[f"energy_{it}"] - kc[f"energy_{it}"]) < 1e-6


def test_periodic_program_on_the_device(libs, tmp_path):
    """tests/cpp/periodic_drop_in.cpp on the device: periodic operators with root and reach, near/far field and precision-tree
    applies through the C++ mirror; node counts identical to the run served by the CPU oracle, values to rounding"""
    src = os.path.join(cb.ROOT, "tests", "cpp", "periodic_drop_in.cpp")
    exe = cb.compile_program([src], str(tmp_path / "periodic_gpu"))
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    kv = cb.key_values(r.stdout)
    check_periodic_values(kv)
    cpu = cb.compile_program([src, os.path.join(cb.ROOT, "tests", "cpp", "oracle_backend.cpp")], str(tmp_path / "periodic_cpu"))
    rc = cb.run_program(cpu, env={"MRCPP_B200_DEVICE": "-1"})
    assert rc.returncode == 0, rc.stderr[-2000:]
    kc = cb.key_values(rc.stdout)
    for key in ("source_nodes", "sol_nodes", "helmholtz_nodes", "scaled_nodes", "poisson_terms", "helmholtz_terms"):
        assert kv[key] == kc[key], key
    for key in ("sol_diff_0_1", "sol_diff_0_h", "near_plus_far", "whole", "helmholtz_sqnorm", "scaled_diff_0_1"):
        assert abs(kv[key] - kc[key]) <= 1e-10 * max(1.0, abs(kc[key])), key
