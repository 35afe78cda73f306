#!/bin/bash
# round 2, GPU call h (8 GPUs): scaling experiment. Default workload (1000 centres) with e2e, 3000 centres resident only, c5, and a
# per-phase profile; host memory of the 8-GPU box.
out=gpurun_out; tag=r02h; mkdir -p $out
free -g > $out/${tag}_host.txt; nproc >> $out/${tag}_host.txt
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err
$TR --master-port 29522 bench.py --gpus 8 --centers 3000 --no-e2e > $out/${tag}_bench_n8_c3000.json 2> $out/${tag}_bench_n8_c3000.err
$TR --master-port 29523 bench.py --gpus 8 --config c5 --no-e2e > $out/${tag}_bench_n8_c5.json 2> $out/${tag}_bench_n8_c5.err
MRX_PROFILE=1 $TR --master-port 29524 tools/scale_probe.py 1000 > $out/${tag}_profile_n8.txt 2>&1
grep "rank 0/8\|device_apply ms\|host phases" $out/${tag}_profile_n8.txt | tail -8
cat $out/${tag}_host.txt
python -c "
import json
for f in ('bench_n8','bench_n8_c3000','bench_n8_c5'):
    try:
        d=json.load(open('$out/${tag}_'+f+'.json')); print(f, d['value'], d['ms_per_step'], d['e2e'], d['detail']['ms_not_in_kernels_per_step'], d['detail']['ms_contract_per_step'], d['detail']['ms_kernel_per_step'], d['detail']['ms_post_per_step'])
    except Exception as e: print(f, 'failed', e)
"
