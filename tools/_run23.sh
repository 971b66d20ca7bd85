cd $GRAFT_REPO_ROOT
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2_bench23.json 2> gpurun_out/r2_bench23.err; cut -c1-200 gpurun_out/r2_bench23.json
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-layers > /dev/null 2> gpurun_out/r2_layers23_unet.txt
timeout 200 python bench.py --workload textseg --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/r2_bench23_textseg.json 2> gpurun_out/r2_layers23_textseg.txt; cut -c1-200 gpurun_out/r2_bench23_textseg.json
timeout 200 python bench.py --workload xception --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/r2_bench23_xception.json 2> gpurun_out/r2_layers23_xception.txt; cut -c1-200 gpurun_out/r2_bench23_xception.json
for w in unet textseg xception; do timeout 200 python tools/kernel_breakdown.py $w > gpurun_out/r2_kb23_$w.txt 2>&1; done; sed -n 3,6p gpurun_out/r2_kb23_unet.txt
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches23.csv python tools/profile_step.py > /dev/null 2>&1; wc -l gpurun_out/r2_launches23.csv
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'pconv_tc|smallco|k2r_|s2d_' -o /tmp/ncu_conv23 -f python tools/profile_step.py > gpurun_out/r2_ncu_conv23.log 2>&1; tail -1 gpurun_out/r2_ncu_conv23.log
ncu -i /tmp/ncu_conv23.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2_ncu_conv23_raw.csv 2>/dev/null; wc -l gpurun_out/r2_ncu_conv23_raw.csv
