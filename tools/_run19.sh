cd $GRAFT_REPO_ROOT
python tools/bench_layer.py pw_512_512 pw_256_256 dec6_192_64 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_pytest19.log 2>&1; tail -2 gpurun_out/r2_pytest19.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-260
python bench.py --workload xception --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-260
