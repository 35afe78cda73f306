#!/bin/bash
# round 2, GPU call l (1 GPU): why the streamed download does not overlap (per-iteration lines, mailbox on/off); transformK pipe with
# the tile-matched warp counts
out=gpurun_out; tag=r02l; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
MRX_PROFILE=1 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_phases.txt 2>&1
MRX_NO_MAILBOX=1 MRX_PROFILE=1 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_phases_nomailbox.txt 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_node_transforms.py -m gpu -q -x -k "bottom_up or top_down or device_projection or node_mw" > $out/${tag}_tests.txt 2>&1
python tools/prof_transform.py 1000 20 5 9 > $out/${tag}_transforms_pipe.txt 2>&1
tail -3 $out/${tag}_tests.txt; cat $out/${tag}_transforms_pipe.txt
grep "mirror True rep 2\|mirror False rep 2" $out/${tag}_e2e_phases.txt $out/${tag}_e2e_phases_nomailbox.txt
