N=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tools/check_multi_gpu.py > gpurun_out/mg_check_$N.log 2>&1; echo "check rc=$?"
grep -v "^\*\|OMP" gpurun_out/mg_check_$N.log | tail -8
BL_GIBBS_TIMING=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 tools/bench_scaling.py > gpurun_out/scaling_$N.log 2> gpurun_out/scaling_$N.err; echo "rc=$?"
grep "^{" gpurun_out/scaling_$N.log | tail -1 > gpurun_out/scaling_$N.json; wc -c gpurun_out/scaling_$N.json; grep "gibbs timing" gpurun_out/scaling_$N.err | sort | uniq -c | sort -rn | head -6
