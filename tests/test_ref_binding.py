"""The reference-side binding of INTEGRATION.md as a real program (tests/cpp/ref_binding.cpp, built by oracle/build_ref.sh against
the reference's own sources compiled in place): MRCPP's FunctionTree<3> / PoissonOperator / HelmholtzOperator are handed to the C
ABI through mrx_tree_from_arrays + mrx_oper_from_arrays (NodeAllocator serial order, OperatorTree::getNode(n, l)), mrx_apply runs,
the result is written back into a reference FunctionTree and compared, INSIDE the reference, with the reference's own
mrcpp::apply. CPU part: the device entry points of the C ABI are served by the oracle (tests/cpp/oracle_backend.cpp, linked into
the _cpu binary only); the GPU part (-m gpu) runs the same program against the product library on the B200."""
import os
import subprocess

import pytest

import cpp_build as cb

REFDIR = os.path.join(cb.ROOT, "oracle", "_ref")


def run_binding(binary, args, device):
    exe = os.path.join(REFDIR, binary)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (oracle/build_ref.sh needs /root/reference at build time)")
    env = dict(os.environ, MWFILTERS_DIR=os.path.join(REFDIR, "mwfilters"), MRX_TABLES=cb.TABLES, MRX_TEST_ORACLE=cb.ORACLE,
               MRCPP_B200_DEVICE=str(device))
    r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, env=env, cwd=cb.ROOT, timeout=1200)
    kv = cb.key_values(r.stdout)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    return kv


def check(kv, prec):
    assert kv["ok"] == 1 and kv["node_set_mismatches"] == 0
    assert kv["g_nodes_ref"] == kv["g_nodes_b200"] > 8
    assert kv["coef_err_floored"] < 1e-12
    print(f"[binding] strict max|d|/||node|| {kv['coef_err_strict']:.3e}, {int(kv['nodes_needing_floor'])} nodes above 1e-12 strictly")
    assert abs(kv["energy_ref"] - kv["energy_b200"]) <= 1e-11 * abs(kv["energy_ref"])
    assert abs(kv["sqnorm_ref"] - kv["sqnorm_b200"]) <= 1e-12 * kv["sqnorm_ref"]
    assert kv["tuples"] > 0 and kv["calc_nodes"] >= kv["g_nodes_b200"]


@pytest.mark.parametrize("kind,order,prec,n", [("poisson", 5, 1e-4, 1), ("helmholtz", 5, 1e-4, 2), ("poisson", 4, 1e-3, 3)])
def test_reference_objects_through_the_c_abi_on_the_oracle_backend(libs, kind, order, prec, n):
    kv = run_binding("ref_binding_cpu", [kind, order, prec, n], -1)
    check(kv, prec)
    if kind == "poisson" and n == 1:  # examples/poisson.cpp's own acceptance check
        assert abs(kv["energy_b200"] - 7.978845608) / 7.978845608 < prec


@pytest.mark.gpu
@pytest.mark.parametrize("kind,order,prec,n", [("poisson", 7, 1e-5, 1), ("poisson", 7, 1e-6, 4), ("helmholtz", 9, 1e-5, 2), ("poisson", 5, 1e-4, 3)])
def test_reference_objects_through_the_c_abi_on_the_device(libs, kind, order, prec, n):
    from mrcpp_b200 import _lib
    if _lib.device() is None or _lib.device() < 0:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    kv = run_binding("ref_binding", [kind, order, prec, n], 0)
    check(kv, prec)
    print("[binding] seconds: reference apply %.3f, binding total %.3f (tree export %.3f, operator export %.3f, apply + download %.3f, "
          "import %.3f); bytes in %d out %d" % (kv["seconds_reference_apply"], kv["seconds_binding_total"], kv["seconds_binding_tree_export"],
                                                kv["seconds_binding_oper_export"], kv["seconds_binding_apply_and_download"],
                                                kv["seconds_binding_result_import"], kv["bytes_in"], kv["bytes_out"]))
