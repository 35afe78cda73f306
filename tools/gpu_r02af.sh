#!/bin/bash
# round 2, GPU call af (1 GPU): ncu --set full of the k = 7 contraction kernel on the final code (iterations 4-6 of the third apply)
out=gpurun_out; tag=r02af; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k "regex:pipe_contract_kernel" -s 22 -c 3 -f -o $out/${tag}_pipe_contract python tools/scale_probe.py 1000 > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_pipe_contract.ncu-rep --page raw --csv > $out/${tag}_pipe_contract_raw.csv 2>/dev/null
ls -la $out/${tag}_pipe_contract*
tail -3 $out/${tag}_ncu.log
