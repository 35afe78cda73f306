#!/bin/bash
# round 2, GPU call k (1 GPU): persistent double-buffered transformK kernel: parity + C2 sweep with and without; e2e phases
out=gpurun_out; tag=r02k; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
MRX_PROFILE=1 python tools/prof_e2e.py 1000 2>&1 | grep -v "iter [0-9]\|host phases\|run_apply_pipe" > $out/${tag}_e2e_phases.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_node_transforms.py tests/test_zz2_gpu_tree_algebra.py -m gpu -q -x -k "bottom_up or top_down or device_projection or node_mw or adaptive or multiply or add" > $out/${tag}_tests.txt 2>&1
python tools/prof_transform.py 1000 20 5 9 > $out/${tag}_transforms_pipe.txt 2>&1
MRX_NO_TPIPE=1 python tools/prof_transform.py 1000 20 5 9 > $out/${tag}_transforms_nopipe.txt 2>&1
tail -3 $out/${tag}_tests.txt; cat $out/${tag}_transforms_pipe.txt; echo ---; cat $out/${tag}_transforms_nopipe.txt; tail -40 $out/${tag}_e2e_phases.txt
