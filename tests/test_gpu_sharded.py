"""Sharded apply on the device (-m gpu). With >= 2 GPUs: torchrun of tools/shard_check.py (every rank must end
with the bit-identical tree of a single-GPU apply). With one GPU: the communicator path with world size 1."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_apply_world1(libs):
    mw, orc = libs
    from mrcpp_b200 import _lib
    if _lib.device() is None or _lib.device() < 0:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    comm = mw.Comm(0, 1, lambda b: b)
    prec = 1e-5
    mra = mw.MultiResolutionAnalysis(7, -4, (-1, -1, -1), (2, 2, 2), 25)
    beta = 100.0
    f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)
    P = mw.PoissonOperator(mra, prec)
    ft = mw.FunctionTree(mra)
    mw.project(prec, ft, f)
    a, b = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sa = mw.apply(prec, a, P, ft)
    sb = mw.apply(prec, b, P, ft, comm=comm)
    A, B = a.to_arrays(), b.to_arrays()
    assert sa.f_applied == sb.f_applied
    assert np.array_equal(A["transl"], B["transl"]) and np.array_equal(A["coefs"], B["coefs"])


@pytest.mark.parametrize("args,env", [(["1e-5", "6"], {}), (["1e-5", "6"], {"MRX_NO_IPC": "1"}), (["1e-4", "4", "11"], {}), (["1e-4", "4", "5"], {}), (["1e-4", "4", "6"], {})])
def test_sharded_apply_two_ranks(libs, args, env):
    """2 ranks: every rank must end with the bit-identical tree of a single-GPU apply (k = 7 with the peer-push and with
    the NCCL coefficient exchange, k = 11, k = 5 and k = 6 (odd K) with the padded contraction kernels)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (covered by tools/shard_check.py under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "shard_check.py")] + args
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("same-topology True") == 2
