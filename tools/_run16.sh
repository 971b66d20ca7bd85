cd $GRAFT_REPO_ROOT
for N in 8 4; do
date +%s > /tmp/t0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench16_n$N.json 2> gpurun_out/r2_bench16_n$N.err; echo "N=$N rc=$? took $(( $(date +%s) - $(cat /tmp/t0) )) s lines=$(wc -l < gpurun_out/r2_bench16_n$N.json)"; cut -c1-250 gpurun_out/r2_bench16_n$N.json; grep -o '"allreduce": "[^"]*"' gpurun_out/r2_bench16_n$N.json; tail -2 gpurun_out/r2_bench16_n$N.err | cut -c1-200
done
