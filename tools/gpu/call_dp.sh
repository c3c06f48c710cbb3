#!/bin/bash
# multi-GPU: bench.py under torchrun (dp_check included), N from $1
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_bench_dp$N.json 2> gpurun_out/r2_bench_dp$N.err
tail -c 1500 gpurun_out/r2_bench_dp$N.json; tail -n 5 gpurun_out/r2_bench_dp$N.err
