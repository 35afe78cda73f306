#!/usr/bin/env python3
"""Pack the multiwavelet filter / cross-correlation tables the hot path needs into one file.

Source: the reference's binary tables `share/mwfilters/I_{H0,G0}_k` (K*K little-endian FP64,
row-major; read by MWFilter::generateBlocks, src/core/MWFilter.cpp:200-251) and
`I_c_{left,right}_k` (K*K x 2K; CrossCorrelation::readCCCBin, src/core/CrossCorrelation.cpp:99-124).
These are numerical DATA (Alpert interpolating multiwavelet filters), not code. They cannot travel
to the GPU box with /root/reference, so the subset needed (interpolating basis only) is packed
into mrcpp_b200/data/mwtables.bin by this script.

Format (little endian): magic 'MRXT', int32 n_entries, then per entry
  int32 kind (0=H0, 1=G0, 2=c_left, 3=c_right, 4/5 = Holoborodko derivative matrices of order 1/2, 6/7/8 = B-spline
  derivative matrices of order 1/2/3), int32 order k, int32 n_doubles, n_doubles * f64.

Kinds 4-8 come from the text tables `I_ph_deriv_{1,2}.txt` / `I_b-spline-deriv{1,2,3}.txt` (per K = k + 1: a line holding K, then
3K rows of K numbers = S_{+1}, S_0, S_{-1}; read by PHCalculator::readSMatrix, src/treebuilders/PHCalculator.cpp:47-79, and
BSCalculator::readSMatrix, BSCalculator.cpp:47-79), orders 1..12.
"""
import struct, sys, os
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/share/mwfilters"
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "mrcpp_b200", "data", "mwtables.bin")

orders = list(range(1, 13))                      # function-tree orders supported
filt_orders = sorted(set(orders) | {2 * k + 1 for k in orders})   # + kernel-tree orders 2k+1
entries = []
for k in filt_orders:
    K = k + 1
    for kind, name in ((0, "I_H0_%d"), (1, "I_G0_%d")):
        a = np.fromfile(os.path.join(src, name % k), dtype="<f8")
        assert a.size == K * K, (name % k, a.size)
        entries.append((kind, k, a))
for k in orders:
    K = k + 1
    for kind, name in ((2, "I_c_left_%d"), (3, "I_c_right_%d")):
        a = np.fromfile(os.path.join(src, name % k), dtype="<f8")
        assert a.size == K * K * 2 * K, (name % k, a.size)
        entries.append((kind, k, a))
for kind, name in ((4, "I_ph_deriv_1.txt"), (5, "I_ph_deriv_2.txt"), (6, "I_b-spline-deriv1.txt"), (7, "I_b-spline-deriv2.txt"),
                   (8, "I_b-spline-deriv3.txt")):
    with open(os.path.join(src, name)) as f:
        lines = f.read().split("\n")
    pos = 0
    for K in range(2, max(orders) + 2):
        assert int(lines[pos].split()[0]) == K, (name, K, lines[pos])
        rows = [[float(x) for x in lines[pos + 1 + i].split()] for i in range(3 * K)]
        assert all(len(r) == K for r in rows), (name, K)
        pos += 1 + 3 * K
        entries.append((kind, K - 1, np.array(rows, dtype="<f8").reshape(-1)))
with open(dst, "wb") as f:
    f.write(b"MRXT")
    f.write(struct.pack("<i", len(entries)))
    for kind, k, a in entries:
        f.write(struct.pack("<iii", kind, k, a.size))
        f.write(a.astype("<f8").tobytes())
print("wrote", dst, os.path.getsize(dst), "bytes,", len(entries), "entries")
