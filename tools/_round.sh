# one GPU call: parity tests, bench (both arms), ncu of the Gibbs kernels
mkdir -p gpurun_out
timeout 480 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r14_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/r14_pytest.log
timeout 150 python bench.py --impl reference > gpurun_out/r14_bench_ref.json 2> gpurun_out/r14_bench_ref.err; echo "ref rc=$?"
timeout 330 python bench.py > gpurun_out/r14_bench.json 2> gpurun_out/r14_bench.err; echo "bench rc=$?"; wc -c gpurun_out/r14_bench.json; tail -3 gpurun_out/r14_bench.err
K='regex:k_xbeta|k_gram_partial|k_gram_reduce|k_devroye|k_beta_draw'
timeout 200 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 12 -c 8 -f -o gpurun_out/prof_r1_14_gibbs python tools/bench_gibbs.py --iters 6 --warmup 2 > gpurun_out/r14_ncu_gibbs.log 2>&1; echo "ncu rc=$?"
