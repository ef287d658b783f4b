mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --deselect tests/test_gpu_big_shapes.py 2>&1 | tail -25 > gpurun_out/r02_fulltests3.txt
for w in C4 C5 C3; do
  timeout 300 python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/r02_n1b_$w.json 2> gpurun_out/r02_n1b_$w.err
done
