#!/bin/bash
# r02ai (2 GPUs): which part of the peer deposit costs the 0.55 ms per step?
TAG=r02ai
mkdir -p gpurun_out
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 $2 2> gpurun_out/${TAG}.err | python -c "
import sys, json
b = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(b['value']), 'ms/step', round(b['ms_per_step'], 3), [round(x, 3) for x in b['ms_per_step_per_rank']], b.get('n_gpu_bit_identical'))" | tee -a gpurun_out/${TAG}_gather_cost.txt
}
run "deposit+commit (default)" ""
MCRT_DIAG_SKIP_COMMIT=1 run "deposit only" ""
MCRT_DIAG_SKIP_DEPOSIT=1 run "commit only" ""
run "contiguous deposit+commit" "--contiguous"
MCRT_DIAG_SKIP_COMMIT=1 run "contiguous deposit only" "--contiguous"
run "no exchange" "--gather none"
tail -2 gpurun_out/${TAG}.err
