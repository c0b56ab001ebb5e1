mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ab_pytest.log
BL_GIBBS_TIMING=1 timeout 300 python tools/bench_models.py --nb-iters 0 --mlogit-iters 10 2>&1 | grep -v "^\*" | tail -3 | cut -c1-420
