mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
OCFFM_PROFILE=2 timeout 400 $TR --master-port 29601 bench.py --gpus 2 --steps 3 --warmup 3 --workload C4 --no-parity-check --no-eval --no-e2e > gpurun_out/r02_n2_C4_phases.json 2> gpurun_out/r02_n2_C4_phases.err
OCFFM_PROFILE=2 timeout 400 python bench.py --steps 3 --warmup 3 --workload C4 --no-cpu-baseline --no-eval --no-e2e > gpurun_out/r02_n1_C4_phases.json 2> gpurun_out/r02_n1_C4_phases.err
OCFFM_PROFILE=2 timeout 400 $TR --master-port 29602 bench.py --gpus 2 --steps 3 --warmup 3 --workload C3 --no-parity-check --no-eval --no-e2e > gpurun_out/r02_n2_C3_phases.json 2> gpurun_out/r02_n2_C3_phases.err
