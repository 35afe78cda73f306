#!/usr/bin/env python3
"""bench.py — Poisson-apply throughput on B200 (BASELINE.json metric).

A "step" is one mrcpp::apply of the 3-D Poisson operator (k=7, prec 1e-7: the north_star target) onto a
synthetic 1000-centre Gaussian density (seed 42, centres uniform in [-8,8]^3, beta log-uniform in
[10,1000], the generator of config C5 of SURVEY.md §8d; 343 K input nodes = 11.2 GB, projected on the device)
with a fresh output tree each step.

  value  output nodes/s (calcNode invocations over all refinement iterations / device time), inputs
         resident in HBM when the timed region starts.
  e2e    same metric through the C-ABI with HOST buffers: the input tree starts in pinned host memory (the apply
         gathers the nodes it reads over PCIe, h2d_bytes_per_step is what actually crossed) and the result tree is
         downloaded inside the timed region.
  roofline  dominant kernel = the contraction kernel (pipe_contract_kernel): algorithmic flops =
         surviving tuples x 6 (k+1)^4, divided by the kernel's CUDA-event time; peak = FP64 tensor
         (DMMA) rate measured in this run (MEASURED_PEAKS.json has no FP64 entry).
  cpu_baseline  the CPU oracle (OpenMP restatement of the reference) on a bounded sample (16 of the 1000 centres).
  --impl reference  the same sample through both CPU implementations: the oracle port and, where oracle/_ref exists, the
         reference's own sources compiled in place; the line carries the faster of the two.

--config selects the workload (default: the headline above): c1 = examples/poisson.cpp at the target precision (k=7, prec 1e-7,
one Gaussian), c4 = HelmholtzOperator(mu=1) k=9 prec 1e-7 on 50 synthetic orbital trees (a step = the 50 applies of one SCF
iteration, examples/scf.cpp:101-111), c5 = Poisson k=11 prec 1e-9 on 100 centres. --centers/--order/--prec override. The
reference arm prints the workload it really times (`config.centers` = the CPU sample, `config.sample_of` = the GPU arm's
workload); the GPU arm additionally times that same CPU-sample workload (`same_workload_as_reference_arm`) so that a ratio on
identical inputs can be formed from the two lines.

N > 1 (torchrun): one process per GPU; ONE apply is sharded over the ranks: every refinement iteration's
output-node list is dealt out cyclically (item i -> rank i % N), input tree and operator are replicated, component norms
are all-gathered over NCCL and the output coefficient blocks are pushed to the peers' HBM by the copy engines over NVLink
(CUDA-IPC mappings; NCCL all-gather if IPC is unavailable) inside the library (mrx_apply_sharded). Same workload for every
N -> strong scaling; time = max over ranks.
"""
import os as _os
import sys as _sys0
if int(_os.environ.get("WORLD_SIZE", "1")) > 1:
    # torchrun pins OMP_NUM_THREADS=1; the host-side input generator (projection) wants the rank's share of cores. The
    # reference arm runs on rank 0 alone (the other ranks exit at once), so there it gets every core of the box.
    _ref_arm = any(a == "reference" or a.endswith("=reference") for a in _sys0.argv)
    _os.environ["OMP_NUM_THREADS"] = str(max(1, (_os.cpu_count() or 1) // (1 if _ref_arm else int(_os.environ["WORLD_SIZE"]))))
# bench prints exactly ONE JSON line on stdout. Libraries write banners there (NCCL prints its version on communicator
# creation on these boxes), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved
# original descriptor.
import sys as _sys
_sys.stdout.flush()
_REAL_STDOUT = _os.dup(1)
_os.dup2(2, 1)


def emit(line):
    _os.write(_REAL_STDOUT, (line + "\n").encode())
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def density(mw, n, seed):
    rng = np.random.default_rng(seed)
    g = mw.GaussExp()
    for _ in range(n):
        beta = 10.0 ** rng.uniform(1, 3)
        g.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-8, 8, 3))))
    return g


class ClockSampler:
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md), read through NVML in a
    background thread (an `nvidia-smi -lms` child process stalls CUDA API calls of the measured process for
    tens of milliseconds per poll, which would corrupt the very numbers it annotates)."""

    def __init__(self, index, period_s=0.02):
        self.index = index
        self.period = period_s
        self.sm = []
        self.reasons = set()
        self.smax = None
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [x for x in vis.split(",") if x != ""]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.ok:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self._stop.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, %d ms period" % int(self.period * 1e3)}


CONFIGS = {
    # name: (order, prec, centres, cpu sample centres, operator)
    "headline": (7, 1e-7, 1000, 16, "poisson"),
    "c1": (7, 1e-7, 1, 1, "poisson"),
    "c4": (9, 1e-7, 50, 1, "helmholtz"),
    "c5": (11, 1e-9, 100, 4, "poisson"),
}


def benzene_orbital(mw, j):
    """C4 input j: 12 benzene-like centres (6 at radius 2.64, 6 at 4.69 bohr, z = 0, 60 degrees apart),
    exponents 1.5 / 0.8, coefficients N(0,1) with seed 2024 + j"""
    rng = np.random.default_rng(2024 + j)
    ge = mw.GaussExp()
    for a in range(12):
        r = 2.64 if a < 6 else 4.69
        ang = math.pi / 3.0 * (a % 6)
        ge.append(mw.GaussFunc(1.5 if a < 6 else 0.8, float(rng.normal()), (r * math.cos(ang), r * math.sin(ang), 0.0)))
    return ge


def workload_inputs(mw, args, centers):
    """the Gaussian expansions of the workload: one density (Poisson configs) or `centers` orbital functions (c4)"""
    if args.operator == "helmholtz":
        return [benzene_orbital(mw, j) for j in range(centers)]
    if args.config == "c1":
        beta = 100.0
        return [[mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)]]
    return [density(mw, centers, 42)]


def make_operator(mw, mra, args):
    return mw.HelmholtzOperator(mra, 1.0, args.prec) if args.operator == "helmholtz" else mw.PoissonOperator(mra, args.prec)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host threads, on a bounded sample of
    the workload: the oracle port (oracle/oracle.cpp: OpenMP restatement, same loop structure and thresholds) and, where
    oracle/_ref was built (oracle/build_ref.sh; the prebuilt files travel to the GPU box), the reference's own sources
    compiled in place. The line reports the faster of the two."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mrcpp_b200 as mw
    from mrcpp_b200 import _lib, build
    build.build_oracle()
    _lib.load().mrx_init(_lib.TABLES.encode(), -1)
    _lib._device = -1
    import oracle_api as orc
    k, prec = args.order, args.prec
    mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
    P = make_operator(mw, mra, args)
    funcs = workload_inputs(mw, args, args.cpu_centers)
    fts = []
    for func in funcs:
        ft = mw.FunctionTree(mra)
        orc.project(prec, ft, func)
        fts.append(ft)
    func = funcs[0]
    times, nodes, tuples = [], 0, 0
    for it in range(args.warmup + args.steps):
        for ft in fts:
            gt = mw.FunctionTree(mra)
            t0 = time.perf_counter()
            st = orc.apply(prec, gt, P, ft)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
                nodes += st.gNodes
                tuples += st.fApplied
    total = sum(times)
    value = nodes / total
    K = k + 1
    port_value = value
    cb = {"value": value, "unit": "nodes/s", "cores": orc.num_threads(), "kind": "port",
          "sample": sample_text(args) + ", full adaptive apply"}
    # the REAL reference, when its sources were compiled in place (oracle/_ref, Eigen replaced by the eager stand-in of
    # oracle/eigen_shim, which costs it temporaries real Eigen does not have): timed on the same sample; the line reports
    # the FASTER of the two CPU implementations so that the ratio against the GPU arm is the conservative one
    import ref_api as ref
    if ref.available():
        try:
            rm = ref.MRA(k, -4, (-1, -1, -1), (2, 2, 2), 25)
            rf = ref.Tree(rm)
            ref.project(prec, rf, list(func))
            RP = ref.helmholtz(rm, 1.0, prec) if args.operator == "helmholtz" else ref.poisson(rm, prec)
            rt, rn = [], 0
            for it in range(1 + min(args.steps, 2)):
                rg = ref.Tree(rm)
                t0 = time.perf_counter()
                ref.apply(prec, rg, RP, rf)
                dt = time.perf_counter() - t0
                if it >= 1:
                    rt.append(dt)
                    rn += rg.n_nodes()
            ref_value = rn / sum(rt)
            cb["port_nodes_per_s"] = port_value
            cb["reference_in_place_nodes_per_s"] = ref_value
            cb["reference_in_place_note"] = ("MRCPP sources compiled in place (oracle/build_ref.sh), dense products through the eager "
                                             "Eigen stand-in; %d threads" % ref.lib().ref_num_threads())
            if ref_value > value:
                value, total = ref_value, nodes / ref_value
                cb["value"], cb["kind"] = ref_value, "reference"
        except Exception as e:  # noqa: BLE001
            cb["reference_in_place_error"] = repr(e)
    line = {
        "impl": "reference", "metric": "helmholtz_apply_output_nodes_per_s" if args.operator == "helmholtz" else "poisson_apply_output_nodes_per_s",
        "value": value, "unit": "nodes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the workload this arm REALLY times: the bounded CPU sample (config.sample_of names the GPU arm's full workload)
        "config": workload_config(args, args.cpu_centers, sample=True),
        "fp64_tflops": tuples * 6 * K ** 4 / (nodes / value) / 1e12,
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def workload_name(args, centers):
    if args.operator == "helmholtz":
        return f"helmholtz_apply_k{args.order}_prec{args.prec:g}_mu1_orbitals{centers}"
    return f"poisson_apply_k{args.order}_prec{args.prec:g}_gauss{centers}"


def sample_text(args):
    if args.operator == "helmholtz":
        return f"the first {args.cpu_centers} of the workload's {args.centers} orbital trees"
    if args.cpu_centers == args.centers:
        return "the whole workload"
    return f"a {args.cpu_centers}-centre density of the workload's generator (same seed; the workload has {args.centers} centres)"


def workload_config(args, centers, sample=False):
    cfg = {"workload": workload_name(args, centers) + ("_cpu_sample" if sample and centers != args.centers else ""),
           "config": args.config}
    if sample and centers != args.centers:
        cfg["sample_of"] = workload_name(args, args.centers)
    cfg.update({
            "order": args.order, "prec": args.prec, "centers": centers, "world": "[-16,16]^3 root scale -4, max depth 25",
            "operator": "HelmholtzOperator(mu=1, prec)" if args.operator == "helmholtz" else "PoissonOperator(prec)",
            "mode": "adaptive (maxIter=-1)",
            "l2_policy": "fresh output tree each step; input tree + operator tables exceed nothing: working set per step "
                         "is re-generated (generated input nodes, output coefficients) and an L2 flush buffer (256 MB) is written between steps",
            "parallelism": "output-node list of every refinement iteration dealt out cyclically over the ranks, input tree + "
                           "operator replicated, NCCL all-gather of component norms, output coefficient blocks pushed to the "
                           "peers over NVLink (CUDA IPC, copy engines)" if args.operator != "helmholtz" else
                           "independent orbital trees dealt out over the ranks (tree j -> rank j % N), no data-path collective"})
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--order", type=int, default=None)
    ap.add_argument("--prec", type=float, default=None)
    ap.add_argument("--centers", type=int, default=None)
    ap.add_argument("--cpu-centers", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm (experiments only: the line then has e2e = null)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    order, prec, centers, cpu_centers, operator = CONFIGS[args.config]
    args.order = order if args.order is None else args.order
    args.prec = prec if args.prec is None else args.prec
    args.centers = centers if args.centers is None else args.centers
    args.cpu_centers = min(cpu_centers if args.cpu_centers is None else args.cpu_centers, args.centers)
    args.operator = operator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import ctypes as C
    import mrcpp_b200 as mw
    from mrcpp_b200 import _lib
    from mrcpp_b200 import build
    if rank == 0 and not os.path.exists(_lib.LIB_PATH):
        build.build_lib()
    if world > 1:
        dist.barrier()
    _lib.init(local_rank)
    L = _lib.load()

    k, prec, K = args.order, args.prec, args.order + 1
    mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
    t0 = time.perf_counter()
    P = make_operator(mw, mra, args)
    t_oper = time.perf_counter() - t0
    by_tree = args.operator == "helmholtz"  # independent orbital trees: dealt out over the ranks, no exchange
    funcs = workload_inputs(mw, args, args.centers)  # same inputs on every rank
    comm = None
    if world > 1 and not by_tree:
        def _bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        comm = mw.Comm(rank, world, _bcast)
    mine = [j for j in range(len(funcs)) if (not by_tree) or j % world == rank]
    t0 = time.perf_counter()
    fts = []
    for j in mine:
        ft = mw.FunctionTree(mra)
        mw.project(prec, ft, funcs[j], device=True)  # quadrature, transforms and norms on the GPU; the host keeps the topology
        ft.sync_device()
        fts.append(ft)
    t_proj = time.perf_counter() - t0

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Acc:
        def __init__(self):
            self.ms = 0.0
            self.nodes = self.tuples = self.tuples_rank = self.launches = self.h2d = self.d2h = self.iters = 0
            self.kern_ms = self.contract_ms = 0.0
            self.phases = {"ms_build": 0.0, "ms_post": 0.0, "ms_upload": 0.0}
            self.last = None

    def one_step(e2e, acc=None, trees=None, oper=None):
        """one pass of the hot path over the workload: one apply per input tree of this rank"""
        trees = fts if trees is None else trees
        oper = P if oper is None else oper
        flush.fill_(1)  # L2 flush between timed iterations (untimed: the timer runs on the library stream)
        torch.cuda.synchronize()
        if e2e:
            for ft in trees:
                ft.drop_device()  # input starts in (pinned) host memory
        step_h2d = step_d2h = 0
        L.mrx_timer_start()
        sts = []
        for ft in trees:
            out = mw.FunctionTree(mra)
            shared = False
            if e2e and comm is not None and shared_arena[0]:
                # sharded apply: the result lands in host memory all ranks have mapped, every rank downloading its share of the
                # chunks over its own PCIe link while the apply runs; the call returns with the whole tree there
                shared = out.set_host_mirror(True, comm=comm)
            elif e2e and (comm is None or rank == 0):
                out.set_host_mirror(True)  # the apply streams the result into host memory while it runs and returns with it there
            st = mw.apply(prec, out, oper, ft, comm=comm)
            if e2e and shared:
                step_d2h += st.d2h_bytes  # this rank's share (summed over ranks below)
            elif e2e and (rank == 0 or by_tree):
                out.sync_host()  # result back in host memory (sharded apply: every rank holds the identical tree, rank 0 reads it back)
                step_d2h += out.nbytes()
            step_h2d += st.h2d_bytes  # counted by the library: coefficient blocks gathered from host memory + norms + topology
            sts.append(st)
            del out
        ms = L.mrx_timer_stop_ms()
        if acc is not None:
            acc.ms += ms
            for st in sts:
                acc.nodes += st.g_nodes
                acc.tuples += st.f_applied
                acc.tuples_rank += st.f_applied_rank
                acc.launches += st.kernel_launches
                acc.iters += st.iterations
                acc.kern_ms += st.ms_kernel
                acc.contract_ms += st.ms_contract
                for kk in acc.phases:
                    acc.phases[kk] += getattr(st, kk)
            acc.last = sts
            acc.h2d, acc.d2h = step_h2d, step_d2h
        return ms

    # ---- resident-input arm
    for _ in range(args.warmup):
        one_step(False)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    A = Acc()
    for _ in range(args.steps):
        one_step(False, A)
    barrier()
    clocks = sampler.stop()

    # ---- end-to-end arm (host buffers in, host buffers out). Every rank keeps its copy of the input in pinned host memory: skipped
    #      (e2e = null, with the reason) when the box does not have the memory for world x input tree
    E = Acc()
    e2e_skip = None
    if args.no_e2e:
        e2e_skip = "--no-e2e"
    else:
        try:
            import psutil
            need = world * sum(ft.nbytes() for ft in fts) * 1.15
            avail = psutil.virtual_memory().available
            if need > avail:
                e2e_skip = "host memory: %d ranks x %.1f GB of pinned input > %.1f GB available" % (world, need / world / 1.15 / 1e9, avail / 1e9)
        except Exception:  # noqa: BLE001
            pass
    if world > 1:
        flag = torch.tensor([1.0 if e2e_skip else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if float(flag[0]) > 0 and not e2e_skip:
            e2e_skip = "skipped on another rank"
    shared_arena = [False]
    if not e2e_skip and comm is not None and not os.environ.get("MRX_BENCH_NO_SHARED_MIRROR"):
        # host arena for the result of one apply (+ slack), mapped by every rank: N PCIe links carry the download instead of one
        out_bytes = max(st.n_nodes_out for st in A.last) * 8 * 8 * (args.order + 1) ** 3
        shared_arena[0] = comm.host_arena(int(out_bytes * 1.05) + (64 << 20))
    if not e2e_skip:
        one_step(True)
        barrier()
        for _ in range(args.steps):
            one_step(True, E)
        barrier()

    tot_ms, e2e_ms = A.ms, E.ms
    nodes, tuples, launches, h2d, d2h = A.nodes, A.tuples, A.launches, E.h2d, E.d2h
    e2e_nodes = E.nodes
    if world > 1:
        t = torch.tensor([tot_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_ms, e2e_ms = float(t[0]), float(t[1])
        # sharded apply: output nodes and surviving tuples are whole-job counts already (every rank holds the full topology; the
        # library sums the tuple counters over ranks). Trees dealt out over the ranks: the counts add up. Launches and copied
        # bytes add up over ranks in both cases.
        c = torch.tensor([launches, h2d, d2h, nodes if by_tree else 0, tuples if by_tree else 0, e2e_nodes if by_tree else 0],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        launches, h2d, d2h = (int(x) for x in c.tolist()[:3])
        if by_tree:
            nodes, tuples, e2e_nodes = (int(x) for x in c.tolist()[3:])

    if rank == 0:
        # FP64 tensor peak measured in this run, with its own clock record (MEASURED_PEAKS.json has no FP64 entry)
        psamp = ClockSampler(local_rank, period_s=0.005)
        psamp.start()
        peak_dmma = L.mrx_bench_dmma_tflops(20000)
        peak_dfma = L.mrx_bench_dfma_tflops(20000)
        peak_clocks = psamp.stop()
        flops = tuples * 6.0 * K ** 4
        # rank 0's contraction kernel and the tuples rank 0 contracted
        achieved = (A.tuples_rank * 6.0 * K ** 4) / (A.contract_ms * 1e-3) / 1e12
        kernel = "pipe_contract_kernel" if k == 7 else "pipe_contract_coop_kernel<%d>" % K
        napply = len(A.last)
        last = A.last[-1]
        line = {
            "metric": "poisson_apply_output_nodes_per_s" if not by_tree else "helmholtz_apply_output_nodes_per_s",
            "value": nodes / (tot_ms * 1e-3), "unit": "nodes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.centers),
            "fp64_tflops": flops / (tot_ms * 1e-3) / 1e12,
            "fp64_tflops_frac_of_dmma_peak": flops / (tot_ms * 1e-3) / 1e12 / (peak_dmma * world),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_dmma, "unit": "TFLOP/s", "frac": achieved / peak_dmma,
                         "traffic": None, "kernel": kernel,
                         "peak_source": "FP64 DMMA m8n8k4 micro-benchmark measured in this run (no FP64 figure in "
                                        "MEASURED_PEAKS.json); DFMA peak %.1f TFLOP/s" % peak_dfma,
                         "peak_clocks": peak_clocks,
                         "peak_arithmetic_limit": "148 SM x 128 flop/clk x sm_mhz",
                         "kernel_share_of_step": A.contract_ms / A.ms, "all_apply_kernels_share_of_step": A.kern_ms / A.ms},
            "clocks": clocks,
            "e2e": ({"value": e2e_nodes / (e2e_ms * 1e-3), "unit": "nodes/s", "h2d_bytes_per_step": int(h2d),
                     "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                     "result_download": ("every rank its chunks, into one host arena mapped by all ranks" if shared_arena[0]
                                         else "rank 0" if world > 1 and not by_tree else "the rank that ran the apply")}
                    if not e2e_skip else None),
            "e2e_skipped": e2e_skip,
            "gpu_launches": int(launches),
            # where the step goes (rank 0's view). Sharded over the ranks: the contraction and the other work-list kernels
            # (screen, scans, fill, reduce). Not sharded or latency-bound: band enumeration launches, the per-iteration norm
            # exchange + split decisions + mailbox reads, and the closing passes every rank repeats on the whole tree
            # (BottomUp, norms, topology to the host; the TopDown(+=) runs inside the loop on a side stream)
            "breakdown_ms": {"contract": A.contract_ms / args.steps, "other_work_list_kernels": (A.kern_ms - A.contract_ms) / args.steps,
                             "closing_passes_replicated": A.phases["ms_post"] / args.steps,
                             "enumeration_exchange_split_host": (A.ms - A.kern_ms - A.phases["ms_post"]) / args.steps,
                             "ms_replicated": (A.ms - A.kern_ms) / args.steps},
            "detail": {"applies_per_step_this_rank": napply, "output_nodes_per_step": nodes // args.steps,
                       "final_tree_nodes_last_apply": last.n_nodes_out, "iterations_last_apply": last.iterations,
                       "tuples_per_step": tuples // args.steps, "generated_input_nodes_last_apply": last.gen_nodes,
                       "input_tree_nodes": sum(ft.getNNodes() for ft in fts),
                       "separation_rank": P.size(), "ms_kernel_per_step": A.kern_ms / args.steps,
                       "ms_contract_per_step": A.contract_ms / args.steps,
                       "ms_build_per_step": A.phases["ms_build"] / args.steps, "ms_post_per_step": A.phases["ms_post"] / args.steps,
                       "ms_not_in_kernels_per_step": (A.ms - A.kern_ms) / args.steps,
                       "setup_s": {"operator": t_oper, "projection": t_proj}},
        }
        # achieved and traffic are both PER LAUNCH averages of the contraction kernel (one launch per refinement iteration)
        line["roofline"]["launches_in_timed_region"] = A.iters
        line["roofline"]["achieved_flop_per_launch"] = A.tuples_rank * 6.0 * K ** 4 / max(A.iters, 1)
        line["roofline"]["traffic"], line["roofline"]["traffic_source"] = ncu_traffic(kernel, k, A.tuples_rank / max(A.iters, 1))
        if world == 1 and args.config == "headline":
            line["transforms"] = transforms_roofline(L, fts[0], K)
        if world == 1 and args.cpu_centers != args.centers and not args.no_e2e:
            # the workload the reference arm times (bench.py --impl reference), on the GPU: a ratio on IDENTICAL inputs can be
            # formed from this object and the reference arm's line. Small workload: launch- and latency-bound on a B200.
            sfuncs = workload_inputs(mw, args, args.cpu_centers)
            sft = []
            for fn in sfuncs:
                t_ = mw.FunctionTree(mra)
                mw.project(prec, t_, fn, device=True)
                t_.sync_device()
                sft.append(t_)
            for _ in range(3):
                one_step(False, None, sft)
            Sa, Se = Acc(), Acc()
            for _ in range(args.steps):
                one_step(False, Sa, sft)
            one_step(True, None, sft)
            for _ in range(args.steps):
                one_step(True, Se, sft)
            line["same_workload_as_reference_arm"] = {
                "config": workload_config(args, args.cpu_centers, sample=True),
                "value": Sa.nodes / (Sa.ms * 1e-3), "unit": "nodes/s", "ms_per_step": Sa.ms / args.steps,
                "e2e_value": Se.nodes / (Se.ms * 1e-3), "e2e_ms_per_step": Se.ms / args.steps,
                "output_nodes_per_step": Sa.nodes // args.steps, "tuples_per_step": Sa.tuples // args.steps}
            # ... and what a binding pays that converts PER CALL instead of keeping handles: the caller's arrays
            # (pageable memory) -> mrx_tree_from_arrays (copy into pinned chunks) -> apply -> mrx_tree_to_arrays; host wall clock
            try:
                arrs = [t_.to_arrays() for t_ in sft]
                tw = []
                for _ in range(args.steps + 1):
                    t0 = time.perf_counter()
                    for a in arrs:
                        fin = mw.FunctionTree.from_arrays(mra, a["scale"], a["transl"], a["parent"], a["child0"], a["coefs"])
                        out = mw.FunctionTree(mra)
                        out.set_host_mirror(True)
                        mw.apply(prec, out, P, fin)
                        res = out.to_arrays()
                        del fin, out
                    tw.append(time.perf_counter() - t0)
                line["same_workload_as_reference_arm"]["from_caller_arrays"] = {
                    "ms_per_step": 1e3 * sum(tw[1:]) / args.steps, "clock": "host wall clock (the copies are host work)",
                    "input_bytes": int(sum(a["coefs"].nbytes for a in arrs)), "output_bytes": int(res["coefs"].nbytes),
                    "path": "mrx_tree_from_arrays + mrx_apply (mirrored output) + mrx_tree_to_arrays, per call"}
                del arrs, res
            except Exception as e:  # noqa: BLE001
                line["same_workload_as_reference_arm"]["from_caller_arrays"] = {"failed": repr(e)[:200]}
            del sft
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is timed at N = 1 only (all host cores free)
            line["cpu_baseline"] = cpu_baseline(args, mw, mra, P)
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic(kernel, order, tuples_per_launch):
    """DRAM traffic of the contraction kernel PER LAUNCH (dram__bytes_read.sum + dram__bytes_write.sum): the bytes-per-tuple figure
    of the committed `ncu --set full` capture (profiles/traffic.json: kernel, order, captured tuples and bytes, source report,
    commit) times the average tuples per launch of THIS run. Returned only when the capture is of the same kernel and order as
    the run; otherwise (None, reason). It is a scaled capture, not a counter read in this run - the source says so."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if not isinstance(t, dict):
            return None, "profiles/traffic.json carries no kernel/order tag"
        if t.get("kernel") != kernel.split("<")[0] or int(t.get("order", -1)) != order:
            return None, "profiles/traffic.json is a capture of %s at k=%s, not of this run's kernel" % (t.get("kernel"), t.get("order"))
        return float(t["bytes_per_tuple"]) * tuples_per_launch, (
            "ncu --set full capture %s (commit %s): %.1f B/tuple x this run's average tuples per launch" %
            (t.get("source"), t.get("commit"), float(t["bytes_per_tuple"])))
    except Exception as e:  # noqa: BLE001
        return None, "no capture (%r)" % (e,)


def transforms_roofline(L, ft, K):
    """C2: filter kernels of mwTransform(TopDown) / (BottomUp) over the bench input tree, CUDA events around the
    level launches only. Algorithmic traffic 128 K^3 B and 96 K^4 flop per parent node: at k=7 that is 6 flop/B, the ridge
    of the FP64 roofline (37 TFLOP/s / 6.5 TB/s = 5.7 flop/B), so both bounds are quoted."""
    import ctypes as C
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    out = {"hbm_peak_gbs": hbm, "hbm_peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}
    nb = C.c_int(0)
    # the += pass last: it accumulates into the tree (scratch by then)
    for kind, name in ((0, "top_down"), (1, "bottom_up"), (2, "top_down_add")):
        L.mrx_bench_mw_transform(ft._h, kind, 2, C.byref(nb))
        ms = L.mrx_bench_mw_transform(ft._h, kind, 20, C.byref(nb))
        per_parent = 192.0 if kind == 2 else 128.0  # += also reads the 8 scaling blocks it adds to
        gbs = nb.value * per_parent * K ** 3 / (ms * 1e-3) / 1e9
        # TopDown with overwrite also zeroes the 7 wavelet blocks of every child (MWNode.cpp:317-319, reference semantics: the
        # children's wavelet coefficients are destroyed): 576 K^3 B per parent really move, 4.5 x the algorithmic 128 K^3 B
        moved = gbs * (4.5 if kind == 0 else 1.0)
        out[name] = {"ms_per_pass": ms, "parent_nodes": nb.value, "nodes_per_s": nb.value / (ms * 1e-3), "achieved_gbs": gbs,
                     "frac_of_hbm_peak": gbs / hbm, "moved_gbs": moved, "moved_frac_of_hbm_peak": moved / hbm,
                     "algorithmic_bytes_per_parent": per_parent * K ** 3,
                     "fp64_tflops": nb.value * 96.0 * K ** 4 / (ms * 1e-3) / 1e12}
    # MWNode::cvTransform (Forward then Backward over every node): element-wise, 128 K^3 B per node and pass
    n_nodes = ft.getNNodes()
    L.mrx_bench_cv_transform(ft._h, 2)
    ms = L.mrx_bench_cv_transform(ft._h, 10)
    gbs = n_nodes * 128.0 * K ** 3 / (ms * 1e-3) / 1e9
    out["cv_transform"] = {"ms_per_pass": ms, "nodes": n_nodes, "nodes_per_s": n_nodes / (ms * 1e-3), "achieved_gbs": gbs,
                           "frac_of_hbm_peak": gbs / hbm, "algorithmic_bytes_per_node": 128.0 * K ** 3}
    return out


def cpu_baseline(args, mw, mra, P):
    """oracle (kind: port) on a bounded sample: same operator, the CPU sample of the workload, all host threads"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mrcpp_b200 import build
    build.build_oracle()
    import oracle_api as orc
    fts = []
    for func in workload_inputs(mw, args, args.cpu_centers):
        ft = mw.FunctionTree(mra)
        orc.project(args.prec, ft, func)
        fts.append(ft)
    # one untimed pass first: in this process host node storage is pinned (cudaMallocHost), and the first pass pays for
    # allocating it; the reference arm (bench.py --impl reference, no CUDA) has no such cost
    orc.apply(args.prec, mw.FunctionTree(mra), P, fts[0])
    nodes = tuples = 0
    dt = 0.0
    for ft in fts:
        gt = mw.FunctionTree(mra)
        t0 = time.perf_counter()
        st = orc.apply(args.prec, gt, P, ft)
        dt += time.perf_counter() - t0
        nodes += st.gNodes
        tuples += st.fApplied
    K = args.order + 1
    return {"value": nodes / dt, "unit": "nodes/s", "cores": orc.num_threads(), "kind": "port",
            "sample": sample_text(args) + f", one full adaptive apply per tree after one warm-up ({dt:.1f} s)",
            "config": workload_config(args, args.cpu_centers, sample=True),
            "fp64_tflops": tuples * 6 * K ** 4 / dt / 1e12, "output_nodes": nodes}


if __name__ == "__main__":
    main()
