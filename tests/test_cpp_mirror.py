"""C++ host mirror of the MRCPP API (include/MRCPP/, header-only over the C ABI), CPU part: the reference's own example
programs compile against it UNMODIFIED, host-side calls agree with the Python mirror of the same C ABI, a hot-path call
without a device aborts like MSG_ABORT, and the program the GPU suite runs (tests/cpp/apply_drop_in.cpp) is checked here
with the device entry points re-defined on top of the CPU oracle (tests/cpp/oracle_backend.cpp, test infrastructure)."""
import math
import os
import signal

import numpy as np
import pytest

import cpp_build as cb

CPP = os.path.join(cb.ROOT, "tests", "cpp")
REF_EXAMPLES = "/root/reference/examples"


@pytest.fixture(scope="module")
def bindir(libs, tmp_path_factory):
    return str(tmp_path_factory.mktemp("cpp"))


def test_host_api_matches_python_mirror(libs, bindir):
    mw, orc = libs
    exe = cb.compile_program([os.path.join(CPP, "host_api.cpp")], os.path.join(bindir, "host_api"))
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "-1"})
    assert r.returncode == 0, r.stderr
    kv = cb.key_values(r.stdout)
    assert (kv["order"], kv["root_scale"], kv["max_scale"], kv["lower"], kv["upper"]) == (7, -4, 21, -16.0, 16.0)
    beta = 100.0
    f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)
    g = mw.GaussFunc(50.0, (50.0 / math.pi) ** 1.5, (0.1, -0.2, 0.3))
    assert abs(kv["self_energy"] - math.sqrt(2 * beta / math.pi)) < 1e-14 * kv["self_energy"]  # GaussFunc.cpp:210-237
    assert abs(kv["pair_energy"] - f.calc_coulomb_energy(g)) < 1e-14
    assert abs(kv["evalf"] - f.evalf((1.0, 1.1, 0.9))) < 1e-14 * abs(kv["evalf"])
    assert kv["exp_size"] == 2 and abs(kv["exp_evalf"] - (f.evalf((0.2, -0.1, 0.4)) + g.evalf((0.2, -0.1, 0.4)))) < 1e-13
    mra = mw.MultiResolutionAnalysis(7, -4, (-1, -1, -1), (2, 2, 2), 25)
    assert kv["poisson_terms"] == mw.PoissonOperator(mra, 1e-5).size() == 73
    assert kv["helmholtz_terms"] == mw.HelmholtzOperator(mra, 1.0, 1e-5).size()
    assert (kv["ph_order"], kv["bs_order"], kv["ph_terms"]) == (2, 3, 1)
    t = mw.FunctionTree(mra)
    mw.build_grid(t, f)
    assert kv["root_nodes"] == 8 and kv["grid_nodes"] == t.getNNodes() and kv["grid_end_nodes"] == t.getNEndNodes()
    t2 = mw.FunctionTree(mra)
    e = mw.GaussExp()
    e.append(f)
    e.append(g)
    mw.build_grid(t2, e)
    assert kv["grid2_nodes"] == kv["copy_nodes"] == t2.getNNodes() and kv["cleared_nodes"] == 8
    assert kv["square_norm_empty"] == -1.0 and kv["log_lines"] == 8
    assert kv["clear_grid_nodes"] == kv["copy_nodes"] and kv["identity_terms"] == 1


def test_hot_path_call_without_device_aborts(libs, bindir):
    """error behaviour of the reference (print + abort, Printer.h:165-169), and no CPU fallback behind the C++ mirror"""
    exe = cb.compile_program([os.path.join(CPP, "apply_drop_in.cpp")], os.path.join(bindir, "apply_drop_in_nodev"))
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "-1"})
    assert r.returncode == -signal.SIGABRT
    assert "no CPU fallback" in r.stderr and "poisson_energy" not in r.stdout


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("example", ["poisson", "projection", "scf", "tree_cleaner", "multiplication"])
def test_reference_examples_compile_unmodified(libs, bindir, example):
    """examples/poisson.cpp, projection.cpp, scf.cpp, tree_cleaner.cpp and multiplication.cpp of the reference, compiled where they lie against
    include/MRCPP/ and linked with libmrcpp_b200.so; run without a device they print their header and abort at the first
    device call"""
    exe = cb.compile_program([os.path.join(REF_EXAMPLES, example + ".cpp")], os.path.join(bindir, "ref_" + example), werror=False)
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "-1"})
    assert r.returncode == -signal.SIGABRT and "no CPU fallback" in r.stderr


def test_drop_in_program_on_the_oracle_backend(libs, bindir):
    """tests/cpp/apply_drop_in.cpp with the device entry points served by the CPU oracle: the reference's known answers, and
    the same numbers as the Python mirror + oracle on the same case"""
    mw, orc = libs
    exe = cb.compile_program([os.path.join(CPP, "apply_drop_in.cpp"), os.path.join(CPP, "oracle_backend.cpp")],
                             os.path.join(bindir, "apply_drop_in_cpu"))
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "-1"})
    assert r.returncode == 0, r.stderr
    kv = cb.key_values(r.stdout)
    assert kv["done"] == 1
    check_drop_in_values(kv)
    # the same Poisson case through the Python mirror and the oracle
    k, prec, beta = 7, 1e-5, 100.0
    mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
    f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)
    ft, gt = mw.FunctionTree(mra), mw.FunctionTree(mra)
    orc.project(prec, ft, f)
    st = orc.apply(prec, gt, mw.PoissonOperator(mra, prec), ft)
    assert kv["poisson_f_nodes"] == ft.getNNodes() and kv["poisson_g_nodes"] == gt.getNNodes()
    assert kv["poisson_tuples"] == st.fApplied and kv["poisson_calc_nodes"] == st.gNodes
    assert abs(kv["poisson_energy"] - orc.dot(gt, ft)) <= 1e-13 * abs(kv["poisson_energy"])
    assert abs(kv["poisson_f_integral"] - ft.integrate()) <= 1e-14 and abs(kv["poisson_g_integral"] - gt.integrate()) <= 1e-10


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="needs the reference tree (build container only)")
def test_reference_scf_example_on_the_oracle_backend(libs, bindir):
    """the reference's examples/scf.cpp, UNMODIFIED, through the C++ mirror with the device entry points served by the CPU oracle:
    the hydrogen SCF must converge to -0.5 Hartree (project, multiply, Helmholtz apply, rescale, add, dot, normalize)"""
    exe = cb.compile_program([os.path.join(REF_EXAMPLES, "scf.cpp"), os.path.join(CPP, "oracle_backend.cpp")], os.path.join(bindir, "ref_scf_cpu"),
                             werror=False)
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "-1"})
    assert r.returncode == 0, r.stderr
    line = [ln for ln in r.stdout.splitlines() if "Eigenvalue" in ln]
    assert line and abs(float(line[0].split()[-1]) + 0.5) < 1e-3


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="needs the reference tree (build container only)")
def test_reference_tree_cleaner_example_on_the_oracle_backend(libs, bindir):
    """the reference's examples/tree_cleaner.cpp, UNMODIFIED (fixed-grid projection, refine_grid, clear_grid until nothing splits):
    the converged function integrates to 1 and has the norm of the projected Gaussian"""
    exe = cb.compile_program([os.path.join(REF_EXAMPLES, "tree_cleaner.cpp"), os.path.join(CPP, "oracle_backend.cpp")],
                             os.path.join(bindir, "ref_tree_cleaner_cpu"), werror=False)
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "-1"})
    assert r.returncode == 0, r.stderr
    vals = {ln.split()[0] + " " + ln.split()[1] if ln.split()[0] == "Square" else ln.split()[0]: float(ln.split()[-1])
            for ln in r.stdout.splitlines() if ln.strip().startswith(("Integral", "Square norm"))}
    assert abs(vals["Integral"] - 1.0) < 1e-6 and abs(vals["Square norm"] - (100.0 / (2 * math.pi)) ** 1.5) < 1e-3


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="needs the reference tree (build container only)")
def test_reference_multiplication_example_on_the_oracle_backend(libs, bindir):
    """the reference's examples/multiplication.cpp, UNMODIFIED (adaptive product with the WaveletAdaptor and with useMaxNorms,
    in-place subtraction of the projected analytic product): integral of the analytic product Gaussian, errors at 1e-9"""
    exe = cb.compile_program([os.path.join(REF_EXAMPLES, "multiplication.cpp"), os.path.join(CPP, "oracle_backend.cpp")],
                             os.path.join(bindir, "ref_multiplication_cpu"), werror=False)
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "-1"})
    assert r.returncode == 0, r.stderr
    vals = {" ".join(ln.split()[:-1]): float(ln.split()[-1]) for ln in r.stdout.splitlines() if "method" in ln}
    beta, d2 = 20.0, 0.27 ** 2
    ana = (beta / math.pi) ** 3 * math.exp(-beta * 0.5 * d2) * (math.pi / (2 * beta)) ** 1.5
    for m in ("1", "2"):
        assert abs(vals["Integral method " + m] - ana) < 1e-6 * ana and vals["Square norm error method " + m] < 1e-6


def test_scf_program_on_the_oracle_backend(libs, bindir):
    """tests/cpp/scf_hydrogen.cpp (the program the GPU suite runs on the device) on the oracle backend: converges to the exact
    ground-state energy, the orbital integrates to 8 sqrt(pi)"""
    exe = cb.compile_program([os.path.join(CPP, "scf_hydrogen.cpp"), os.path.join(CPP, "oracle_backend.cpp")], os.path.join(bindir, "scf_cpu"))
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "-1"})
    assert r.returncode == 0, r.stderr
    check_scf_values(cb.key_values(r.stdout))


def check_scf_values(kv):
    assert kv["done"] == 1 and 3 <= kv["iterations"] <= 12
    assert abs(kv["final_energy"] + 0.5) < 1e-4 and kv["final_update"] < 1e-3
    assert abs(kv["orbital_norm"] - 1.0) < 1e-12 and abs(kv["orbital_integral"] - 8.0 * math.sqrt(math.pi)) < 0.05


def check_drop_in_values(kv, what="all"):
    """known answers of tests/cpp/apply_drop_in.cpp (shared with the GPU tests); what: 'core', 'algebra' or 'all'"""
    prec = 1e-5
    if what != "algebra":
        assert kv["poisson_terms"] == 73
        assert abs(kv["poisson_analytic"] - 7.978845608) < 1e-9
        assert abs(kv["poisson_energy"] - kv["poisson_analytic"]) / kv["poisson_analytic"] < prec   # tests/operators/poisson_operator.cpp
        assert abs(kv["poisson_f_integral"] - 1.0) < 1e-9
        assert kv["poisson_fixed_grid_nodes"] == kv["poisson_g_nodes"]
        assert abs(kv["poisson_fixed_grid_energy"] - kv["poisson_analytic"]) / kv["poisson_analytic"] < prec
        assert kv["poisson_tuples"] > 1e5 and kv["poisson_calc_nodes"] >= kv["poisson_g_nodes"]
        # tests/operators/helmholtz_operator.cpp: norm and overlap of the fixed point within apply_prec
        assert abs(kv["helmholtz_out_norm"] - 1.0) < 3e-2 and abs(kv["helmholtz_overlap"] - 1.0) < 3e-2
        for d in range(3):
            assert kv[f"derivative_{d}_rel_err"] < 1e-4 and abs(kv[f"derivative_{d}_sqnorm"] - kv["derivative_0_sqnorm"]) < 1e-6
    if what != "core":
        # gradient / divergence / add: div grad f against the analytic Laplacian; <div grad f | f> = -|grad f|^2; integral 0
        assert kv["divergence_rel_err"] < 1e-3 and abs(kv["divergence_integral"]) < 1e-8
        assert abs(kv["divergence_overlap"] + kv["divergence_grad_sqnorm"]) < 1e-9 * kv["divergence_grad_sqnorm"]
        assert kv["gradient_point_rel_err"] < 1e-3 and kv["function_point_rel_err"] < 1e-4  # derivative_operator.cpp:417-453
        # adaptive add (examples/addition.cpp): integral 1 - 2 + 3, point value, linearity of the overlap
        assert kv["addition_nodes"] > 8 and abs(kv["addition_integral"] - 2.0) < 1e-6 and kv["addition_point_rel_err"] < 1e-2
        assert abs(kv["addition_overlap_1"] - kv["addition_overlap_expected"]) < 1e-3 * abs(kv["addition_overlap_expected"])
        # multiply / square (examples/multiplication.cpp): product of two Gaussians against the analytic product Gaussian
        assert kv["multiplication_nodes"] > 8 and kv["multiplication_rel_err"] < 1e-2 and kv["multiplication_point_rel_err"] < 1e-2
        assert abs(kv["multiplication_integral"] - kv["multiplication_integral_analytic"]) < 1e-3 * kv["multiplication_integral_analytic"]
        assert abs(kv["square_integral"] - kv["square_expected"]) < 1e-3 * kv["square_expected"]
        assert abs(kv["dot_vectors_integral"] - kv["dot_vectors_expected"]) < 1e-3 * abs(kv["dot_vectors_expected"])
    if what == "all":
        assert abs(kv["divergence_grad_sqnorm"] - 3.0 * kv["derivative_0_sqnorm"]) < 1e-6


def check_periodic_values(kv):
    """numbers tests/cpp/periodic_drop_in.cpp prints: the periodic Poisson solution against the analytic one (the point
    (1,0,0) is the image of (-1,0,0): evalf wraps as periodic::coord_manipulation does), near + far field = whole, the
    Helmholtz apply, and the precision-tree apply"""
    assert kv["periodic"] == 1 and kv["done"] == 1
    assert kv["poisson_terms"] > 10 and kv["helmholtz_terms"] > 10
    assert kv["source_nodes"] > 8 and kv["sol_nodes"] > 8
    assert abs(kv["sol_diff_0_1"] - kv["exact_diff_0_1"]) < 2e-3
    assert abs(kv["sol_diff_0_h"] - kv["exact_diff_0_h"]) < 2e-3
    assert abs(kv["near_plus_far"] - kv["whole"]) < 1e-5 * abs(kv["whole"]) + 1e-6
    assert kv["helmholtz_nodes"] >= 8 and kv["helmholtz_sqnorm"] > 0
    assert kv["scaled_nodes"] >= 8 and abs(kv["scaled_diff_0_1"] - kv["exact_diff_0_1"]) < 2e-3


def test_periodic_program_on_the_oracle_backend(libs, bindir):
    """tests/cpp/periodic_drop_in.cpp (periodic BoundingBox, operators with root and reach, apply / apply_near_field /
    apply_far_field, apply with precision trees -- the calls of the reference's periodic Poisson/Helmholtz tests) runs on CPU
    with the device entry points served by the oracle"""
    exe = cb.compile_program([os.path.join(CPP, "periodic_drop_in.cpp"), os.path.join(CPP, "oracle_backend.cpp")],
                             os.path.join(bindir, "periodic_cpu"))
    r = cb.run_program(exe, env={"MRCPP_B200_DEVICE": "-1"})
    assert r.returncode == 0, r.stderr[-2000:]
    check_periodic_values(cb.key_values(r.stdout))
