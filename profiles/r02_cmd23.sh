mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_schedule_knobs.py tests/test_gpu_edge_cases.py tests/test_gpu_parity.py tests/test_train_cli.py tests/test_reference_main.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_tests_final.txt
timeout 300 python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_n1c_C3.json 2> gpurun_out/r02_n1c_C3.err
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_n1c_C2.json 2> gpurun_out/r02_n1c_C2.err
