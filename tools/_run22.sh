cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q -x --timeout=120 --timeout-method=thread > gpurun_out/r2_pytest22.log 2>&1; tail -3 gpurun_out/r2_pytest22.log; grep -E "^FAILED|^ERROR|Timeout" gpurun_out/r2_pytest22.log | head
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-260
timeout 120 python bench.py --workload xception --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-260
