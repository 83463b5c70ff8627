cd /root/repo
mkdir -p gpurun_out
timeout 120 ncu --set full --clock-control none -k regex:gemm_kernel --launch-skip 10 --launch-count 5 \
  -o gpurun_out/r2b_tc_gemm python tools/ncu_tc_linear.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r2b_tc_gemm.ncu-rep --page raw --csv > gpurun_out/r2b_tc_gemm_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_tc_gemm.ncu-rep --page details > gpurun_out/r2b_tc_gemm_ncu_details.txt 2>/dev/null
wc -l gpurun_out/r2b_tc_gemm_ncu_raw.csv
