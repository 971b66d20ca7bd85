cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest18.log 2>&1; tail -3 gpurun_out/r2_pytest18.log; grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest18.log | head
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-layers > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err; cut -c1-300 gpurun_out/r2_bench18.json
python bench.py --workload textseg --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/r2_bench18_textseg.json 2> gpurun_out/r2_bench18_textseg.err; cut -c1-300 gpurun_out/r2_bench18_textseg.json
python bench.py --workload xception --steps 10 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/r2_bench18_xception.json 2> gpurun_out/r2_bench18_xception.err; cut -c1-300 gpurun_out/r2_bench18_xception.json
python tools/bench_layer.py pw_512_512 pw_256_256 dec6_192_64 2>&1 | tail -6
