#!/bin/bash
# round 2, GPU call t (1 GPU): end-to-end path -- gather of a host-resident input beside the contraction (arrival flags), mirrored
# output of a large iteration in node sub-ranges; tests of both, bench, e2e timings with each piece switched off
out=gpurun_out; tag=r02t; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mirror or host_memory or end_to_end or full_size" > $out/${tag}_tests.txt 2>&1
tail -3 $out/${tag}_tests.txt
timeout 300 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python -c "
import json
d=json.load(open('$out/${tag}_bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['breakdown_ms'])
"
for v in default nooverlap nosub neither sub8 sub2; do
  case $v in
    default) env_="";;
    nooverlap) env_="MRX_NO_FETCH_OVERLAP=1";;
    nosub) env_="MRX_SUB_RANGES=1";;
    neither) env_="MRX_NO_FETCH_OVERLAP=1 MRX_SUB_RANGES=1";;
    sub8) env_="MRX_SUB_RANGES=8 MRX_SUB_MIN_TILES=2";;
    sub2) env_="MRX_SUB_RANGES=2";;
  esac
  env $env_ MRX_E2E_MIRROR_ONLY=1 timeout 200 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_$v.txt 2>&1
  echo "$v: $(grep 'mirror True' $out/${tag}_e2e_$v.txt | tail -2 | tr '\n' ' ')"
done
MRX_PROFILE=1 MRX_E2E_MIRROR_ONLY=1 timeout 200 python tools/prof_e2e.py 1000 > $out/${tag}_e2e_phases.txt 2>&1
grep "device_apply ms\|run_apply_pipe ms\|push of the\|drained" $out/${tag}_e2e_phases.txt | tail -8
