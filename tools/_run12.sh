cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest12.log 2>&1; tail -4 gpurun_out/r2_pytest12.log; grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest12.log | head
python bench.py --steps 20 --warmup 5 --profile-layers > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; cut -c1-330 gpurun_out/r2_bench12.json
PCB_DISABLE_S2D_STEM=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-330
python tools/kernel_breakdown.py > gpurun_out/r2_kb12.txt 2>&1; sed -n 3,50p gpurun_out/r2_kb12.txt
python bench.py --workload textseg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench12_textseg.json 2> /dev/null; cut -c1-330 gpurun_out/r2_bench12_textseg.json
