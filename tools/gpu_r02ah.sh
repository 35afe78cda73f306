#!/bin/bash
# round 2, GPU call ah (1 GPU): 16-warp CTA of the persistent transform kernel at K = 10 (ncu of the 8-warp CTA: DMMA 50 %, 2 warps per
# scheduler): transform tests with it, C2 sweep at k = 9 with 16 and with 8 warps
out=gpurun_out; tag=r02ah; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
MRX_TPIPE_WARPS=16 timeout 70 python -m pytest tests/test_gpu_parity.py tests/test_gpu_node_transforms.py -m gpu -q -x -k "bottom_up or top_down or node_mw or baseline" > $out/${tag}_tests_w16.txt 2>&1
tail -2 $out/${tag}_tests_w16.txt
MRX_TPIPE_WARPS=16 timeout 35 python tools/prof_transform.py 1000 20 9 > $out/${tag}_transforms_k9_w16.txt 2>&1
timeout 35 python tools/prof_transform.py 1000 20 9 > $out/${tag}_transforms_k9_w8.txt 2>&1
echo w16; cat $out/${tag}_transforms_k9_w16.txt; echo w8; cat $out/${tag}_transforms_k9_w8.txt
