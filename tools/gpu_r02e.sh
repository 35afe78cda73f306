#!/bin/bash
# round 2, GPU call e: M-stacked first stage of the K = 6 / 10 / 12 contraction: parity for every order, bench c5 / c4 with and without
out=gpurun_out; tag=r02e; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "adaptive or variants or fixed_grid or c4 or c5 or helmholtz or linearity" > $out/${tag}_tests.txt 2>&1
python bench.py --config c5 --steps 3 --no-cpu-baseline > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
MRX_NO_STACK=1 python bench.py --config c5 --steps 3 --no-cpu-baseline > $out/${tag}_bench_c5_nostack.json 2> $out/${tag}_bench_c5_nostack.err
python bench.py --config c4 --steps 3 --no-cpu-baseline > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
MRX_NO_STACK=1 python bench.py --config c4 --steps 3 --no-cpu-baseline > $out/${tag}_bench_c4_nostack.json 2> $out/${tag}_bench_c4_nostack.err
tail -5 $out/${tag}_tests.txt
python -c "
import json
for f in ('bench_c5','bench_c5_nostack','bench_c4','bench_c4_nostack'):
    d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['detail']['ms_contract_per_step'])
"
