mkdir -p gpurun_out
timeout 200 python profiles/e2e_timing.py > gpurun_out/r02_e2e_timing.txt 2>&1
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_f.json 2> gpurun_out/r02_bench_c2_f.err
# classic (per-iteration) kernels for the per-kernel DRAM / L1 / tensor numbers: one warm outer iteration first
export OCFFM_PERSIST_CG=0 OCFFM_MROW=0
timeout 400 ncu --set full --clock-control none -k 'regex:k_gram_tc|k_rowgemm_tc|k_grad_cross|k_sddmm_add|k_spmm_update|k_hess_cross|k_colsum_w' -s 20 -c 22 -f -o gpurun_out/r02c_C2 python profiles/one_epoch.py C2 32 1 > gpurun_out/r02_ncu_f_C2.log 2>&1
ncu -i gpurun_out/r02c_C2.ncu-rep --page raw --csv > gpurun_out/r02c_C2_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k 'regex:k_gram_tc|k_rowgemm_tc' -c 4 -f -o gpurun_out/r02c_C4 python profiles/one_epoch.py C4 32 0 > gpurun_out/r02_ncu_f_C4.log 2>&1
ncu -i gpurun_out/r02c_C4.ncu-rep --page raw --csv > gpurun_out/r02c_C4_raw.csv 2>/dev/null
unset OCFFM_PERSIST_CG OCFFM_MROW
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02c_launches_C4.csv python profiles/one_epoch.py C4 32 1 > gpurun_out/r02_ncu_l_C4.log 2>&1
du -sh gpurun_out; ls -la gpurun_out
sz=$(du -sm gpurun_out | cut -f1); if [ "$sz" -gt 55 ]; then rm -f gpurun_out/*.ncu-rep; fi
