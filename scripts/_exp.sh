mkdir -p gpurun_out
timeout -k 5 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "steady/" -k regex:"dgemm_nt|tc_gemm" -c 7 -f -o gpurun_out/r02s_solver python scripts/ncu_targets.py --K 100000 > gpurun_out/r02s_ncu.log 2>&1
tail -3 gpurun_out/r02s_ncu.log; ls -la gpurun_out/*.ncu-rep
