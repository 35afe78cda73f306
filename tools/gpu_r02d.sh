#!/bin/bash
# round 2, GPU call d (2 GPUs): sharded apply with the side-stream unpack: bit-identity tests, bench at N=2 and N=1, profile
out=gpurun_out; tag=r02d; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.txt 2>&1
python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > $out/${tag}_tests_sharded.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
python bench.py --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
MRX_PROFILE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/scale_probe.py 1000 > $out/${tag}_profile_n2.txt 2>&1
tail -5 $out/${tag}_tests_sharded.txt; tail -6 $out/${tag}_profile_n2.txt
python -c "
import json
for f in ('${tag}_bench_n1.json','${tag}_bench_n2.json'):
    d=json.load(open('$out/'+f)); print(f, d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['detail']['ms_not_in_kernels_per_step'], d['detail']['ms_kernel_per_step'], d['detail']['ms_post_per_step'])
"
