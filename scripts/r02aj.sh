#!/bin/bash
# r02aj (2 GPUs): direct device output (post kernel writes into the caller's buffer / the interleaved slots) -- tests, N=1 and N=2 bench
TAG=r02aj
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
for d in 1 0; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --option direct_out=$d 2> gpurun_out/${TAG}_n1.err | python -c "
import sys, json
b = json.loads(sys.stdin.read())
print('N=1 direct_out=$d', round(b['value']), 'ms/step', round(b['ms_per_step'], 3), {k: round(v, 3) for k, v in b['roofline']['stage_ms'].items()}, 'e2e', round(b['e2e']['value']), 'sweep', round(b['sweep_workload_1gpu']['value']))" | tee -a gpurun_out/${TAG}_direct_out.txt
done
for g in p2p none; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --gather $g 2> gpurun_out/${TAG}_n2.err | tee gpurun_out/${TAG}_bench_n2_$g.json | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2 gather=$g', round(b['value']), 'ms/step', round(b['ms_per_step'], 3), [round(x, 3) for x in b['ms_per_step_per_rank']], b.get('n_gpu_bit_identical'), 'e2e', round(b['e2e']['value']))" | tee -a gpurun_out/${TAG}_direct_out.txt
done
tail -3 gpurun_out/${TAG}_n2.err
