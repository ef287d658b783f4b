mkdir -p gpurun_out
B="python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu-baseline --no-eval --no-e2e"
OCFFM_PROFILE=2 timeout 200 $B > gpurun_out/r02_c3diag_default.json 2> gpurun_out/r02_c3diag_default.err
OCFFM_PROFILE=2 OCFFM_MROW=0 timeout 200 $B > gpurun_out/r02_c3diag_mrow0.json 2> gpurun_out/r02_c3diag_mrow0.err
OCFFM_PROFILE=2 OCFFM_PERSIST_CG=0 timeout 200 $B > gpurun_out/r02_c3diag_persist0.json 2> gpurun_out/r02_c3diag_persist0.err
