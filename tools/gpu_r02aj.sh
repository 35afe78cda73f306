#!/bin/bash
# round 2, GPU call aj (1 GPU), last: transform tests, baseline-config tests and smoke on the final library (16-warp persistent transform CTA at K = 10)
out=gpurun_out; tag=r02aj; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 75 python -m pytest tests/test_gpu_parity.py tests/test_gpu_node_transforms.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "bottom_up or top_down or node_mw or baseline or adaptive or end_to_end" > $out/${tag}_tests.txt 2>&1
tail -2 $out/${tag}_tests.txt
timeout 30 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.txt 2>&1
tail -1 $out/${tag}_smoke.txt
timeout 25 python tools/prof_transform.py 1000 20 9 5 > $out/${tag}_transforms_k9_k5.txt 2>&1; cat $out/${tag}_transforms_k9_k5.txt
