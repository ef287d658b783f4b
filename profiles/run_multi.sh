#!/bin/bash
# N-GPU bench lines of round 2 (run under `gpurun --gpus N`): profiles/run_multi.sh N "C2 C4 C3 C5" [tests]
N=$1; WL=${2:-"C2 C4 C3 C5"}; TESTS=$3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_n${N}_tests.txt; fi
port=29511
for w in $WL; do
  port=$((port+1))
  extra="--no-parity-check"; steps=3
  if [ "$w" = "C2" ]; then extra=""; steps=5; fi
  timeout 600 $TR --master-port $port bench.py --gpus $N --steps $steps --warmup 3 --workload $w $extra \
      > gpurun_out/r02_n${N}_${w}.json 2> gpurun_out/r02_n${N}_${w}.err
  echo "$w rc=$?" >> gpurun_out/r02_n${N}_rc.txt
done
nvidia-smi --query-gpu=index,name,memory.used --format=csv > gpurun_out/r02_n${N}_smi.txt
