cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest10.log 2>&1; tail -4 gpurun_out/r2_pytest10.log
python bench.py --steps 20 --warmup 5 --profile-layers > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; cut -c1-330 gpurun_out/r2_bench10.json
python tools/kernel_breakdown.py > gpurun_out/r2_kb10.txt 2>&1; sed -n 3,12p gpurun_out/r2_kb10.txt
python bench.py --workload textseg --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/r2_bench10_textseg.json 2> gpurun_out/r2_bench10_textseg.err; cut -c1-330 gpurun_out/r2_bench10_textseg.json
python bench.py --workload xception --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/r2_bench10_xception.json 2> gpurun_out/r2_bench10_xception.err; cut -c1-330 gpurun_out/r2_bench10_xception.json
python tools/kernel_breakdown.py textseg > gpurun_out/r2_kb10_textseg.txt 2>&1; sed -n 3,12p gpurun_out/r2_kb10_textseg.txt
python tools/kernel_breakdown.py xception > gpurun_out/r2_kb10_xception.txt 2>&1; sed -n 3,12p gpurun_out/r2_kb10_xception.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches10.csv python tools/profile_step.py > /dev/null 2>&1; wc -l gpurun_out/r2_launches10.csv
timeout 1500 ncu --profile-from-start off --set full --clock-control none -k regex:'pconv_tc|smallco|k2r_' -o gpurun_out/r2_ncu_conv10 -f python tools/profile_step.py > gpurun_out/r2_ncu_conv10.log 2>&1; tail -2 gpurun_out/r2_ncu_conv10.log; ls -la gpurun_out/r2_ncu_conv10.ncu-rep
