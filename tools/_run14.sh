cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-260
PCB_TMA_L2_PREFETCH=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-layers 2> gpurun_out/r2_bench14_pf.err | cut -c1-260
PCB_TMA_L2_PREFETCH=1 python bench.py --workload xception --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-260
PCB_TMA_L2_PREFETCH=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tma or lc_ or network_bf16" 2>&1 | tail -2
