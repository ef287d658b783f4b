#!/bin/bash
# Round-2 ncu evidence (run under gpurun, ONE GPU).  Every program first runs without ncu.
set -x
mkdir -p gpurun_out
K='regex:k_gram_tc|k_rowgemm_tc|k_colsum_w|k_spmm_rows|k_spmm_update|k_grad_cross|k_sddmm_add|k_hess_cross|k_hess_heavy|k_row_gram|k_rowgemm|k_cg_step|k_side_rows|k_side_diag_iter'
python profiles/one_epoch.py C2 32 2 > gpurun_out/r02_one_C2.txt 2>&1 || exit 1
# launch list of one warm outer iteration + the bench command itself
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02b_launches_C2.csv \
    python profiles/one_epoch.py C2 32 2 > gpurun_out/r02_ncu_l_C2.log 2>&1
# full sets: skip the setup + first two iterations (about 2 x 600 + 150 launches), then 260 solver launches
ncu --set full --clock-control none --import-source on -k "$K" -s 700 -c 160 -f -o gpurun_out/r02b_solver_C2 \
    python profiles/one_epoch.py C2 32 2 > gpurun_out/r02_ncu_f_C2.log 2>&1
python profiles/one_epoch.py C4 32 1 > gpurun_out/r02_one_C4b.txt 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02b_launches_C4.csv \
    python profiles/one_epoch.py C4 32 1 > gpurun_out/r02_ncu_l_C4.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_gram_tc|k_rowgemm_tc|k_hess_cross|k_grad_cross|k_sddmm_add|k_spmm_update|k_side_rows' -s 900 -c 40 -f \
    -o gpurun_out/r02b_solver_C4 python profiles/one_epoch.py C4 32 1 > gpurun_out/r02_ncu_f_C4.log 2>&1
ls -la gpurun_out/*.ncu-rep
