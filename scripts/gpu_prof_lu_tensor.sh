# full ncu captures of the tensor-LU kernels at N = 1024 (batch 296 = two matrices per SM): panel 0 of the panel kernel, the
# long-K update kernels of panel 4 (column form, then row form), the L assembly and the solve kernel
mkdir -p gpurun_out
CMD="python scripts/quick_time.py lu 1024 296"
timeout 300 $CMD > gpurun_out/plain_lut.log 2>&1 || exit 1
cat gpurun_out/plain_lut.log
ncu --set full --clock-control none --import-source on -k regex:lu_cpanel -s 0 -c 1 -f -o gpurun_out/prof_lut_panel $CMD > gpurun_out/ncu_lut_panel.log 2>&1; tail -1 gpurun_out/ncu_lut_panel.log
ncu --set full --clock-control none --import-source on -k regex:lu_ll_kernel -s 7 -c 2 -f -o gpurun_out/prof_lut_ll $CMD > gpurun_out/ncu_lut_ll.log 2>&1; tail -1 gpurun_out/ncu_lut_ll.log
ncu --set full --clock-control none --import-source on -k regex:lu_assemble -s 0 -c 1 -f -o gpurun_out/prof_lut_asm $CMD > gpurun_out/ncu_lut_asm.log 2>&1; tail -1 gpurun_out/ncu_lut_asm.log
ncu --set full --clock-control none --import-source on -k regex:lu_trsm -s 2 -c 1 -f -o gpurun_out/prof_lut_trsm $CMD > gpurun_out/ncu_lut_trsm.log 2>&1; tail -1 gpurun_out/ncu_lut_trsm.log
