python -m pytest tests/test_gpu_edge_cases.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_t6.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_d.json 2> gpurun_out/r02_bench_c2_d.err
OCFFM_MROW=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02b_launches_C2_mrow2.csv python profiles/one_epoch.py C2 32 2 > gpurun_out/r02_ncu_l_C2m.log 2>&1
bash profiles/ncu_capture.sh > gpurun_out/r02_ncu_capture.log 2>&1
