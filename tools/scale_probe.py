"""Sharded-apply timing probe (torchrun or plain python): ms per apply for several densities, no tree downloads."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import mrcpp_b200 as mw
from mrcpp_b200 import _lib
_lib.init(lr)
comm = None
if world > 1:
    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    comm = mw.Comm(rank, world, bcast)
k = int(os.environ.get("MRX_PROBE_K", "7")); prec = float(os.environ.get("MRX_PROBE_PREC", "1e-7"))
mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
P = mw.PoissonOperator(mra, prec)
for n in [int(a) for a in sys.argv[1:]] or [100]:
    rng = np.random.default_rng(42)
    func = mw.GaussExp()
    for i in range(n):
        beta = 10.0 ** rng.uniform(1, 3)
        func.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-8, 8, 3))))
    f = mw.FunctionTree(mra); mw.project(prec, f, func, device=True)
    ts = []
    for rep in range(6):
        g = mw.FunctionTree(mra)
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t = time.time()
        st = mw.apply(prec, g, P, f, comm=comm)
        torch.cuda.synchronize(); ts.append((time.time() - t) * 1e3)
        del g
    if os.environ.get("MRX_PROFILE") is None or rank == 0:
        print(f"rank {rank}/{world} centres {n}: apply ms {['%.1f' % x for x in ts]} contract {st.ms_contract:.1f} kernel {st.ms_kernel:.1f} post {st.ms_post:.1f} "
              f"nodes {st.g_nodes} tuples {st.f_applied}", flush=True)
    del f
if world > 1:
    dist.barrier(); dist.destroy_process_group()
