cd $GRAFT_REPO_ROOT
date +%s > /tmp/t0
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q -k "two_rank" > gpurun_out/r2_pytest15_2rank.log 2>&1; tail -3 gpurun_out/r2_pytest15_2rank.log
echo "tests took $(( $(date +%s) - $(cat /tmp/t0) )) s"; date +%s > /tmp/t0
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench15_n2.json 2> gpurun_out/r2_bench15_n2.err; echo "bench rc=$? took $(( $(date +%s) - $(cat /tmp/t0) )) s"; wc -l gpurun_out/r2_bench15_n2.json; cut -c1-300 gpurun_out/r2_bench15_n2.json; tail -3 gpurun_out/r2_bench15_n2.err
date +%s > /tmp/t0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench15_ref_n2.json 2>/dev/null; echo "ref rc=$? took $(( $(date +%s) - $(cat /tmp/t0) )) s"; cut -c1-200 gpurun_out/r2_bench15_ref_n2.json
