mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_schedule_knobs.py tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r02_t7.txt
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_e.json 2> gpurun_out/r02_bench_c2_e.err
OCFFM_PERSIST_CG=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eval --no-e2e > gpurun_out/r02_bench_c2_p0.json 2> gpurun_out/r02_bench_c2_p0.err
OCFFM_PERSIST_CG=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eval --no-e2e > gpurun_out/r02_bench_c2_p1.json 2> gpurun_out/r02_bench_c2_p1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02c_launches_C2.csv python profiles/one_epoch.py C2 32 2 > gpurun_out/r02_ncu_l_C2.log 2>&1
