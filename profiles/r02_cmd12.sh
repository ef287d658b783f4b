mkdir -p gpurun_out
timeout 200 python profiles/e2e_timing.py > gpurun_out/r02_e2e_timing2.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_n1_C2.json 2> gpurun_out/r02_n1_C2.err
for w in C4 C3 C5; do
  timeout 400 python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/r02_n1_$w.json 2> gpurun_out/r02_n1_$w.err
done
