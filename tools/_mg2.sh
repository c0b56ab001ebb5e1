# Multi-GPU round: parity check, stage timing of the sharded logit sweep (both kernels), scaling bench.
#   gpurun --gpus N --timeout 900 -- 'bash tools/_mg2.sh N TAG'
mkdir -p gpurun_out
N=${1:-2}; T=${2:-mg}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29555 tools/check_multi_gpu.py > gpurun_out/${T}_check_$N.log 2>&1; echo "check rc=$?"; grep -v "^\*\|OMP" gpurun_out/${T}_check_$N.log | tail -8
BL_GIBBS_TIMING=1 timeout 120 $RUN --master-port 29556 tools/bench_gibbs.py --iters 200 > gpurun_out/${T}_gibbs_$N.log 2>&1; grep "timing\|iters_per_sec" gpurun_out/${T}_gibbs_$N.log | tail -3
BL_GIBBS_TIMING=1 timeout 120 $RUN --master-port 29557 tools/bench_gibbs.py --iters 200 --one-pass > gpurun_out/${T}_gibbs_onepass_$N.log 2>&1; grep "timing\|iters_per_sec" gpurun_out/${T}_gibbs_onepass_$N.log | tail -3
timeout 120 $RUN --master-port 29558 tools/bench_gibbs.py --iters 40 --constrained > gpurun_out/${T}_gibbs_constrained_$N.log 2>&1; grep "iters_per_sec" gpurun_out/${T}_gibbs_constrained_$N.log | tail -1
