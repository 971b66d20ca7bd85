cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest13.log 2>&1; tail -3 gpurun_out/r2_pytest13.log; grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest13.log | head
python bench.py --steps 20 --warmup 5 --profile-layers > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; cut -c1-330 gpurun_out/r2_bench13.json
python tools/kernel_breakdown.py > gpurun_out/r2_kb13.txt 2>&1; sed -n 3,12p gpurun_out/r2_kb13.txt; grep -n "s2d\|wgrad_tma_kernel<64" gpurun_out/r2_kb13.txt
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'pconv_tc_tma_kernel|dw4_s1_kernel|dw4_s1_wgrad' --launch-skip 60 -c 5 -o gpurun_out/r2_ncu_xc13 -f python tools/profile_step.py xception > gpurun_out/r2_ncu_xc13.log 2>&1; tail -2 gpurun_out/r2_ncu_xc13.log; ls -la gpurun_out/*.ncu-rep
