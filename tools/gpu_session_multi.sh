#!/bin/bash
# Multi-GPU GPU call (gpurun --gpus N): the global-solve bench line and the coefficient-by-coefficient check.
# usage: bash tools/gpu_session_multi.sh <tag> <ngpus>
TAG=${1:-dev}; N=${2:-2}; WHAT=${3:-all}
mkdir -p gpurun_out
if [ "$WHAT" = all ]; then python -m pytest tests/test_gpu_api_surface.py -m gpu -q -k follow 2>&1 | tail -3 > gpurun_out/${TAG}_device_test.log; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
tail -c 1200 gpurun_out/${TAG}_bench_${N}gpu.json; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
if [ "$WHAT" = all ]; then timeout 600 $TR tools/check_global_solve.py > gpurun_out/${TAG}_check_${N}gpu.log 2>&1; tail -5 gpurun_out/${TAG}_check_${N}gpu.log; fi
du -sh gpurun_out
