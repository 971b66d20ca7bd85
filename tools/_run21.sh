cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest21.log 2>&1; tail -3 gpurun_out/r2_pytest21.log; grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest21.log | head
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-260
python bench.py --workload textseg --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-260
python bench.py --workload xception --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-260
