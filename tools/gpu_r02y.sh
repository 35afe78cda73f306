#!/bin/bash
# round 2, GPU call y (1 GPU): look-ahead flag load without a compiler barrier (calls u-x: the waiting contraction kernel ran 6 % below the plain one whatever the block loads were)
# kernel, one CTA per SM); tests, e2e phases, bench
out=gpurun_out; tag=r02y; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mirror or host_memory or end_to_end or full_size" > $out/${tag}_tests.txt 2>&1
tail -3 $out/${tag}_tests.txt
for v in default nooverlap; do
  case $v in
    default) env_="";;
    nooverlap) env_="MRX_NO_FETCH_OVERLAP=1";;
  esac
  env $env_ MRX_PROFILE=1 MRX_E2E_MIRROR_ONLY=1 MRX_E2E_KEEP=1 timeout 100 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_$v.txt 2>&1
  echo "$v: $(grep 'mirror True\|partially' $out/${tag}_e2e_$v.txt | tail -4 | tr '\n' ' ')"
  grep "iter [4-7] nG" $out/${tag}_e2e_$v.txt | sed -n 9,12p | cut -c1-150
  grep "device_apply ms\|run_apply_pipe ms\|push of the" $out/${tag}_e2e_$v.txt | sed -n 13,15p
done
timeout 300 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python -c "
import json
d=json.load(open('$out/${tag}_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['breakdown_ms'])
"
