cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest11.log 2>&1; tail -4 gpurun_out/r2_pytest11.log
python bench.py --steps 20 --warmup 5 --profile-layers > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; cut -c1-330 gpurun_out/r2_bench11.json
python tools/kernel_breakdown.py > gpurun_out/r2_kb11.txt 2>&1; sed -n 3,8p gpurun_out/r2_kb11.txt
python bench.py --workload textseg --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/r2_bench11_textseg.json 2> gpurun_out/r2_bench11_textseg.err; cut -c1-330 gpurun_out/r2_bench11_textseg.json
python bench.py --workload xception --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/r2_bench11_xception.json 2> gpurun_out/r2_bench11_xception.err; cut -c1-330 gpurun_out/r2_bench11_xception.json
python tools/kernel_breakdown.py textseg > gpurun_out/r2_kb11_textseg.txt 2>&1; sed -n 3,12p gpurun_out/r2_kb11_textseg.txt
python tools/kernel_breakdown.py xception > gpurun_out/r2_kb11_xception.txt 2>&1; sed -n 3,12p gpurun_out/r2_kb11_xception.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches11.csv python tools/profile_step.py > /dev/null 2>&1; wc -l gpurun_out/r2_launches11.csv
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subunit_op_utcmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'pconv_tc|smallco|k2r_' -o /tmp/ncu_conv11 -f python tools/profile_step.py > gpurun_out/r2_ncu_conv11.log 2>&1; tail -2 gpurun_out/r2_ncu_conv11.log
ncu -i /tmp/ncu_conv11.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2_ncu_conv11_raw.csv 2>/dev/null; wc -l gpurun_out/r2_ncu_conv11_raw.csv
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'k2r_combine|k2r_dbuild|renorm_bwd_vec|scse_bwd' -c 6 -o gpurun_out/r2_ncu_k2r11 -f python tools/profile_step.py > /dev/null 2>&1; ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
