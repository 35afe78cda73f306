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
    P = mw.PoissonOperator(mra, prec)
    func = density(mw, args.cpu_centers, 42)
    ft = mw.FunctionTree(mra)
    orc.project(prec, ft, func)
    times, nodes, tuples = [], 0, 0
    for it in range(args.warmup + args.steps):
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
          "sample": f"{args.cpu_centers}-centre subset of the workload density, full adaptive apply"}
    # the REAL reference, when its sources were compiled in place (oracle/_ref, Eigen replaced by the eager stand-in of
    # oracle/eigen_shim, which costs it temporaries real Eigen does not have): timed on the same sample; the line reports
    # the FASTER of the two CPU implementations so that the ratio against the GPU arm is the conservative one
    import ref_api as ref
    if ref.available():
        try:
            rm = ref.MRA(k, -4, (-1, -1, -1), (2, 2, 2), 25)
            rf = ref.Tree(rm)
            ref.project(prec, rf, list(func))
            RP = ref.poisson(rm, prec)
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
        "impl": "reference", "metric": "poisson_apply_output_nodes_per_s", "value": value, "unit": "nodes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.centers),  # the GPU arm's workload; each step here is the bounded sample below
        "fp64_tflops": tuples * 6 * K ** 4 / (nodes / value) / 1e12,
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def workload_config(args, centers, sample=False):
    return {"workload": f"poisson_apply_k{args.order}_prec{args.prec:g}_gauss{centers}" + ("_cpu_sample" if sample else ""),
            "order": args.order, "prec": args.prec, "centers": centers, "world": "[-16,16]^3 root scale -4, max depth 25",
            "operator": "PoissonOperator(prec)", "mode": "adaptive (maxIter=-1)",
            "l2_policy": "fresh output tree each step; input tree + operator tables exceed nothing: working set per step "
                         "is re-generated (generated input nodes, output coefficients) and an L2 flush buffer (256 MB) is written between steps",
            "parallelism": "output-node list of every refinement iteration dealt out cyclically over the ranks, input tree + "
                           "operator replicated, NCCL all-gather of component norms, output coefficient blocks pushed to the "
                           "peers over NVLink (CUDA IPC, copy engines)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--order", type=int, default=7)
    ap.add_argument("--prec", type=float, default=1e-7)
    ap.add_argument("--centers", type=int, default=1000)
    ap.add_argument("--cpu-centers", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

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
    P = mw.PoissonOperator(mra, prec)
    t_oper = time.perf_counter() - t0
    func = density(mw, args.centers, 42)  # same density on every rank: the apply is sharded, not the data
    comm = None
    if world > 1:
        def _bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        comm = mw.Comm(rank, world, _bcast)
    ft = mw.FunctionTree(mra)
    t0 = time.perf_counter()
    mw.project(prec, ft, func, device=True)  # quadrature, transforms and norms on the GPU; the host keeps the topology
    t_proj = time.perf_counter() - t0
    ft.sync_device()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(e2e):
        flush.fill_(1)  # L2 flush between timed iterations (untimed: the timer runs on the library stream)
        torch.cuda.synchronize()
        if e2e:
            ft.drop_device()  # input starts in (pinned) host memory
        out = mw.FunctionTree(mra)
        L.mrx_timer_start()
        st = mw.apply(prec, out, P, ft, comm=comm)
        if e2e and rank == 0:
            out.sync_host()  # result back in host memory (every rank holds the identical tree in HBM; rank 0 reads it back)
        ms = L.mrx_timer_stop_ms()
        nbytes_out = out.nbytes()
        del out
        return st, ms, nbytes_out

    # ---- resident-input arm
    for _ in range(args.warmup):
        one_step(False)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    tot_ms = 0.0
    nodes = tuples = launches = 0
    kern_ms = 0.0
    contract_ms = 0.0
    phases = {"ms_build": 0.0, "ms_post": 0.0, "ms_upload": 0.0}
    last = None
    for _ in range(args.steps):
        st, ms, _ = one_step(False)
        tot_ms += ms
        nodes += st.g_nodes
        tuples += st.f_applied
        launches += st.kernel_launches
        kern_ms += st.ms_kernel
        contract_ms += st.ms_contract
        for kk in phases:
            phases[kk] += getattr(st, kk)
        last = st
    barrier()
    clocks = sampler.stop()

    # ---- end-to-end arm (host buffers in, host buffers out)
    one_step(True)
    barrier()
    e2e_ms = 0.0
    e2e_nodes = 0
    h2d = d2h = 0
    for _ in range(args.steps):
        st, ms, nb_out = one_step(True)
        e2e_ms += ms
        e2e_nodes += st.g_nodes
        h2d = st.h2d_bytes  # counted by the library: coefficient blocks gathered from host memory + norms + topology
        d2h = nb_out if rank == 0 else 0
    barrier()

    if world > 1:
        t = torch.tensor([tot_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_ms, e2e_ms = float(t[0]), float(t[1])
        # output nodes and surviving tuples are whole-job counts already (every rank holds the full topology; the library
        # sums the tuple counters over ranks); launches and copied bytes add up over ranks
        c = torch.tensor([launches, h2d, d2h], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        launches, h2d, d2h = (int(x) for x in c.tolist())
        km = torch.tensor([kern_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(km, op=dist.ReduceOp.MAX)
        kern_ms_max = float(km[0])
    else:
        kern_ms_max = kern_ms

    if rank == 0:
        peak_dmma = L.mrx_bench_dmma_tflops(20000)
        peak_dfma = L.mrx_bench_dfma_tflops(20000)
        flops = tuples * 6.0 * K ** 4
        # rank 0's contraction kernel and the tuples rank 0 contracted
        achieved = (last.f_applied_rank * 6.0 * K ** 4 * args.steps) / (contract_ms * 1e-3) / 1e12
        line = {
            "metric": "poisson_apply_output_nodes_per_s", "value": nodes / (tot_ms * 1e-3), "unit": "nodes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.centers),
            "fp64_tflops": flops / (tot_ms * 1e-3) / 1e12,
            "fp64_tflops_frac_of_dmma_peak": flops / (tot_ms * 1e-3) / 1e12 / (peak_dmma * world),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_dmma, "unit": "TFLOP/s", "frac": achieved / peak_dmma,
                         "traffic": None, "kernel": "pipe_contract_kernel" if k == 7 else "pipe_contract_coop_kernel",
                         "peak_source": "FP64 DMMA m8n8k4 micro-benchmark measured in this run (no FP64 figure in "
                                        "MEASURED_PEAKS.json); DFMA peak %.1f TFLOP/s" % peak_dfma,
                         "kernel_share_of_step": contract_ms / tot_ms, "all_apply_kernels_share_of_step": kern_ms / tot_ms},
            "clocks": clocks,
            "e2e": {"value": e2e_nodes / (e2e_ms * 1e-3), "unit": "nodes/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "detail": {"output_nodes_per_step": last.g_nodes, "final_tree_nodes": last.n_nodes_out, "iterations": last.iterations,
                       "tuples_per_step": last.f_applied, "generated_input_nodes": last.gen_nodes, "input_tree_nodes": ft.getNNodes(),
                       "separation_rank": P.size(), "ms_kernel_per_step": kern_ms / args.steps, "ms_contract_per_step": contract_ms / args.steps,
                       "ms_build_per_step": phases["ms_build"] / args.steps, "ms_post_per_step": phases["ms_post"] / args.steps,
                       "setup_s": {"operator": t_oper, "projection": t_proj}},
        }
        line["roofline"]["traffic"] = ncu_traffic()
        line["transforms"] = transforms_roofline(L, ft, K)
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is timed at N = 1 only (all host cores free)
            line["cpu_baseline"] = cpu_baseline(args, mw, mra, P)
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the contraction kernel from the committed ncu --set full capture
    (profiles/traffic.json, written by tools/ncu_summary.py): bytes of the largest captured launch, or None"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return None


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
    for kind, name in ((0, "top_down"), (1, "bottom_up")):
        L.mrx_bench_mw_transform(ft._h, kind, 2, C.byref(nb))
        ms = L.mrx_bench_mw_transform(ft._h, kind, 20, C.byref(nb))
        gbs = nb.value * 128.0 * K ** 3 / (ms * 1e-3) / 1e9
        # TopDown with overwrite also zeroes the 7 wavelet blocks of every child (MWNode.cpp:317-319): 576 K^3 B per parent
        # really move, 4.5 x the algorithmic 128 K^3 B
        moved = gbs * (4.5 if kind == 0 else 1.0)
        out[name] = {"ms_per_pass": ms, "parent_nodes": nb.value, "nodes_per_s": nb.value / (ms * 1e-3), "achieved_gbs": gbs,
                     "frac_of_hbm_peak": gbs / hbm, "moved_gbs": moved, "moved_frac_of_hbm_peak": moved / hbm,
                     "fp64_tflops": nb.value * 96.0 * K ** 4 / (ms * 1e-3) / 1e12}
    return out


def cpu_baseline(args, mw, mra, P):
    """oracle (kind: port) on a bounded sample: same operator, cpu_centers-centre density, all host threads"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mrcpp_b200 import build
    build.build_oracle()
    import oracle_api as orc
    func = density(mw, args.cpu_centers, 42)
    ft = mw.FunctionTree(mra)
    orc.project(args.prec, ft, func)
    # one untimed pass first: in this process host node storage is pinned (cudaMallocHost), and the first pass pays for
    # allocating it; the reference arm (bench.py --impl reference, no CUDA) has no such cost
    orc.apply(args.prec, mw.FunctionTree(mra), P, ft)
    gt = mw.FunctionTree(mra)
    t0 = time.perf_counter()
    st = orc.apply(args.prec, gt, P, ft)
    dt = time.perf_counter() - t0
    K = args.order + 1
    return {"value": st.gNodes / dt, "unit": "nodes/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"{args.cpu_centers}-centre subset of the workload density, one full adaptive apply after one warm-up ({dt:.1f} s)",
            "fp64_tflops": st.fApplied * 6 * K ** 4 / dt / 1e12, "output_nodes": st.gNodes}


if __name__ == "__main__":
    main()
