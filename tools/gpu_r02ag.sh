#!/bin/bash
# round 2, GPU call ag (1 GPU): ncu --set full of the persistent double-buffered transform kernel (transformK_pipe_kernel) at k = 9 and k = 5
out=gpurun_out; tag=r02ag; mkdir -p $out
timeout 170 ncu --set full --clock-control none --import-source on -k "regex:transformK_pipe" -c 16 -f -o $out/${tag}_transformK_pipe python tools/prof_transform.py 300 2 9 5 > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_transformK_pipe.ncu-rep --page raw --csv > $out/${tag}_transformK_pipe_raw.csv 2>/dev/null
sz=$(stat -c %s $out/${tag}_transformK_pipe.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 20000000 ]; then rm -f $out/${tag}_transformK_pipe.ncu-rep; fi
ls -la $out/${tag}_*; tail -4 $out/${tag}_ncu.log
