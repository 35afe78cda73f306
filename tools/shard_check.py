"""2+ GPU check of the sharded apply (torchrun): every rank must end with the tree a single-GPU apply gives."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import mrcpp_b200 as mw
from mrcpp_b200 import _lib
_lib.init(lr)

def bcast(b):
    obj = [b]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]

comm = mw.Comm(rank, world, bcast)
prec = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-5; n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
k = int(sys.argv[3]) if len(sys.argv) > 3 else 7
mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
rng = np.random.default_rng(42)
func = mw.GaussExp()
for i in range(n):
    beta = 10.0 ** rng.uniform(1, 3)
    func.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-8, 8, 3))))
P = mw.PoissonOperator(mra, prec)
f = mw.FunctionTree(mra); mw.project(prec, f, func)
for rep in range(3):
    ref = mw.FunctionTree(mra); torch.cuda.synchronize(); t = time.time(); s1 = mw.apply(prec, ref, P, f); t1 = time.time() - t
for rep in range(3):
    g = mw.FunctionTree(mra)
    dist.barrier(); torch.cuda.synchronize(); t = time.time()
    s2 = mw.apply(prec, g, P, f, comm=comm)
    dt = time.time() - t
# the same sharded apply with the input tree in pinned host memory: every rank gathers the nodes ITS items read
f.drop_device()
g2 = mw.FunctionTree(mra)
dist.barrier()
s3 = mw.apply(prec, g2, P, f, comm=comm)
# result into host memory shared by the ranks: every rank downloads its chunks, all see the whole tree when the call returns
A, B = g.to_arrays(), ref.to_arrays()
have_arena = comm.host_arena(int(ref.nbytes() * 1.1) + (64 << 20))
if os.environ.get("MRX_EXPECT_ARENA"):
    assert have_arena, "shared host arena expected on this box"
for rep in range(2):  # second round: arena space of the first tree came back
    g3 = mw.FunctionTree(mra)
    shared = g3.set_host_mirror(True, comm=comm)
    assert shared == have_arena
    dist.barrier()
    s4 = mw.apply(prec, g3, P, f, comm=comm)
    g3.drop_device()  # what to_arrays reads now is the host copy
    C3 = g3.to_arrays()
    assert np.array_equal(C3["transl"], B["transl"]) and np.array_equal(C3["coefs"], B["coefs"]), "shared host mirror differs"
    if shared:
        t = torch.tensor([float(s4.d2h_bytes)], dtype=torch.float64, device="cuda"); dist.all_reduce(t)
        assert int(t[0]) == g3.nbytes(), (int(t[0]), g3.nbytes())
        assert 0 < s4.d2h_bytes < g3.nbytes()
    del g3
print(f"rank {rank}: shared host mirror {'on' if have_arena else 'unavailable'} ok", flush=True)
C2 = g2.to_arrays()
assert np.array_equal(C2["transl"], B["transl"]) and np.array_equal(C2["coefs"], B["coefs"]), "sharded apply on a host-resident input differs"
assert 0 < s3.h2d_bytes < f.nbytes() and s3.f_applied == s1.f_applied
same = np.array_equal(A["scale"], B["scale"]) and np.array_equal(A["transl"], B["transl"])
err = np.abs(A["coefs"] - B["coefs"]).max() if same else -1
print(f"rank {rank}/{world}: nodes {len(A['scale'])} same-topology {same} max|dcoef| {err:.3e} tuples sharded {s2.f_applied} single {s1.f_applied} "
      f"sharded apply {dt*1e3:.1f} ms (single {t1*1e3:.1f} ms, contract {s1.ms_contract:.1f}) contract {s2.ms_contract:.1f} ms", flush=True)
assert same and err == 0.0 and s2.f_applied == s1.f_applied
dist.barrier()
dist.destroy_process_group()
