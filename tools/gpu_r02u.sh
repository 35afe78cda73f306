#!/bin/bash
# round 2, GPU call u (1 GPU): gather beside the contraction with a RELAXED flag load + L2-only block loads (the acquire load of
# call t invalidated L1 at every source switch: overlap cost 7 ms); same tests and e2e variants
out=gpurun_out; tag=r02u; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mirror or host_memory or end_to_end or full_size" > $out/${tag}_tests.txt 2>&1
tail -3 $out/${tag}_tests.txt
for v in default nooverlap nosub neither; do
  case $v in
    default) env_="";;
    nooverlap) env_="MRX_NO_FETCH_OVERLAP=1";;
    nosub) env_="MRX_SUB_RANGES=1";;
    neither) env_="MRX_NO_FETCH_OVERLAP=1 MRX_SUB_RANGES=1";;
  esac
  env $env_ MRX_E2E_MIRROR_ONLY=1 timeout 200 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_$v.txt 2>&1
  echo "$v: $(grep 'mirror True' $out/${tag}_e2e_$v.txt | tail -2 | tr '\n' ' ')"
done
MRX_PROFILE=1 MRX_E2E_MIRROR_ONLY=1 timeout 200 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_phases.txt 2>&1
grep "device_apply ms\|run_apply_pipe ms\|push of the\|drained" $out/${tag}_e2e_phases.txt | tail -6
grep "iter [4-7] " $out/${tag}_e2e_phases.txt | tail -4
timeout 300 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python -c "
import json
d=json.load(open('$out/${tag}_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['breakdown_ms'])
"
