mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_n2b_tests.txt
timeout 300 $TR --master-port 29701 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_n2b_C2.json 2> gpurun_out/r02_n2b_C2.err
echo "C2 rc=$?" >> gpurun_out/r02_n2b_rc.txt
timeout 300 $TR --master-port 29702 bench.py --gpus 2 --steps 3 --warmup 3 --workload C4 --no-parity-check > gpurun_out/r02_n2b_C4.json 2> gpurun_out/r02_n2b_C4.err
echo "C4 rc=$?" >> gpurun_out/r02_n2b_rc.txt
timeout 300 $TR --master-port 29703 bench.py --gpus 2 --steps 3 --warmup 3 --workload C5 --no-parity-check > gpurun_out/r02_n2b_C5.json 2> gpurun_out/r02_n2b_C5.err
echo "C5 rc=$?" >> gpurun_out/r02_n2b_rc.txt
