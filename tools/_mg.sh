mkdir -p gpurun_out
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tools/check_multi_gpu.py > gpurun_out/mg_check_$N.log 2>&1; echo "check rc=$?"; grep -v "^\*\|OMP" gpurun_out/mg_check_$N.log | tail -6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; wc -c gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
