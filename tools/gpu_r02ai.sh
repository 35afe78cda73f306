#!/bin/bash
# round 2, GPU call ai (1 GPU): warps per CTA of the persistent transform kernel at K = 10: 16 (new default), 17, 25; tests with the default
out=gpurun_out; tag=r02ai; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_gpu_node_transforms.py -m gpu -q -x -k "bottom_up or top_down or node_mw or baseline" > $out/${tag}_tests.txt 2>&1
tail -2 $out/${tag}_tests.txt
for w in 16 17 25; do
MRX_TPIPE_WARPS=$w timeout 30 python tools/prof_transform.py 1000 20 9 > $out/${tag}_transforms_k9_w$w.txt 2>&1
echo w$w; cat $out/${tag}_transforms_k9_w$w.txt
done
