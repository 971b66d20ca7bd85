cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q -x -k "lazycat or lc_" > gpurun_out/r2_pytest9_lc.log 2>&1; tail -5 gpurun_out/r2_pytest9_lc.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest9.log 2>&1; tail -15 gpurun_out/r2_pytest9.log
python bench.py --steps 20 --warmup 5 --profile-layers > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; cut -c1-400 gpurun_out/r2_bench9.json
PCB_DISABLE_K2R=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-330
python tools/kernel_breakdown.py > gpurun_out/r2_kb9.txt 2>&1; head -30 gpurun_out/r2_kb9.txt
python bench.py --workload textseg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench9_textseg.json 2> gpurun_out/r2_bench9_textseg.err; cut -c1-330 gpurun_out/r2_bench9_textseg.json
python bench.py --workload xception --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench9_xception.json 2> gpurun_out/r2_bench9_xception.err; cut -c1-330 gpurun_out/r2_bench9_xception.json
python tools/kernel_breakdown.py textseg > gpurun_out/r2_kb9_textseg.txt 2>&1; head -14 gpurun_out/r2_kb9_textseg.txt
