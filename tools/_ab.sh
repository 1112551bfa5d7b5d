T=tools/time_stream_kernels.py
FDM_RHS_NOPAIRS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"assemble_rhs_batch" -c 2 -o gpurun_out/r03_rhs -f python $T > gpurun_out/ncu_asm.log 2>&1
tail -3 gpurun_out/ncu_asm.log
