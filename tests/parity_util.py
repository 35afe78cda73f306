"""Coefficient-parity measure shared by the parity tests.

north_star bar: per-node FP64 coefficients within 1e-12 RELATIVE TO THE NODE NORM. Two figures are computed for every
comparison and both are reported (printed, and appended to gpurun_out/parity_report.jsonl when that directory exists so
that the GPU run leaves a record):

  strict   max over nodes of max|a - b| / ||b_node||, no floor at all (nodes whose reference norm is exactly 0 are
           compared absolutely against the largest node norm).
  floored  the same, but a node whose norm is below `floor` x the largest node norm is measured against that floor:
           the rounding of an operator application is relative to the INPUT neighbourhood (~1e-16 x |O| x |f|), so an
           output node of norm 1e-14 x the largest cannot agree to 1e-12 of ITSELF between two summation orders; the
           real reference and the oracle (same algorithm, different summation order in the dense products) show
           exactly the same effect.

The assertion is on the floored figure; the report carries the strict one, the number of nodes whose strict figure
exceeds the tolerance (= the nodes that needed the floor) and the largest relative norm among those nodes, so the
relaxation is visible instead of silent.
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def coef_parity(Acoefs, Bcoefs, tol=1e-12, floor=1e-3, label=None):
    """A, B: [nNodes][8 K^3] coefficient arrays of two trees with the SAME node order; B is the reference side.
    Returns a dict with the strict and floored figures; prints one line."""
    A = np.asarray(Acoefs)
    B = np.asarray(Bcoefs)
    assert A.shape == B.shape, (A.shape, B.shape)
    if A.size == 0:
        return {"nodes": 0, "strict": 0.0, "floored": 0.0, "needed_floor": 0, "largest_rel_norm_needing_floor": 0.0}
    nrm = np.sqrt((B * B).sum(axis=1))
    err = np.abs(A - B).max(axis=1)
    nmax = float(nrm.max())
    strict_den = np.where(nrm > 0.0, nrm, nmax + 1e-300)
    strict = err / strict_den
    floored = err / np.maximum(nrm, floor * nmax + 1e-300)
    need = strict > tol
    rep = {
        "label": label or os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0],
        "nodes": int(len(nrm)),
        "tol": tol,
        "floor": floor,
        "strict": float(strict.max()),
        "strict_p99": float(np.quantile(strict, 0.99)),
        "floored": float(floored.max()),
        "needed_floor": int(need.sum()),
        "largest_rel_norm_needing_floor": float((nrm[need] / (nmax + 1e-300)).max()) if need.any() else 0.0,
        "max_abs_err_over_max_norm": float(err.max() / (nmax + 1e-300)),
    }
    print("[parity] %(label)s: nodes %(nodes)d strict max|d|/||node|| %(strict).3e (p99 %(strict_p99).3e) floored(%(floor)g) "
          "%(floored).3e; %(needed_floor)d nodes above %(tol)g strictly, their largest ||node||/max||node|| = "
          "%(largest_rel_norm_needing_floor).2e" % rep)
    if os.path.isdir(os.path.dirname(REPORT)):
        try:
            with open(REPORT, "a") as f:
                f.write(json.dumps(rep) + "\n")
        except OSError:
            pass
    return rep
