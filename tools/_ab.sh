mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ab_pytest.log
for P in 128 256; do echo "P=$P"; BL_GIBBS_TIMING=1 timeout 120 python tools/bench_gibbs.py --iters 40 --P $P 2>&1 | grep "timing" | tail -1; done
BL_GIBBS_TIMING=1 timeout 300 python tools/bench_models.py --nb-iters 6 --mlogit-iters 0 2>&1 | grep -v "^\*" | tail -2 | cut -c1-300
