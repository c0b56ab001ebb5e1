# Profiling round (one GPU): ncu --set full captures of the kernels bench.py quotes, on the current tree.
#   gpurun --timeout 1500 -- 'bash tools/_prof.sh r2'
# Every ncu run follows a plain run of the same command that exited 0 (no pipe after it).
mkdir -p gpurun_out
T=${1:-rX}
NCU="ncu --set full --clock-control none --import-source on -f"
B1="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-e2e"
$B1 > gpurun_out/${T}_prof_plain1.log 2>&1 && $NCU -k regex:"k_sp_setup|k_loop_regroup" -c 4 -o gpurun_out/prof_${T}_sp $B1 > gpurun_out/${T}_prof_ncu1.log 2>&1; echo "sp rc=$?"
B2="python bench.py --workload pg1 --draws 134217728 --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-e2e"
$B2 > gpurun_out/${T}_prof_plain2.log 2>&1 && $NCU -k regex:k_devroye_refill -s 1 -c 1 -o gpurun_out/prof_${T}_pg1 $B2 > gpurun_out/${T}_prof_ncu2.log 2>&1; echo "pg1 rc=$?"
B3="python tools/bench_gibbs.py --iters 6 --warmup 2 --two-pass"
$B3 > gpurun_out/${T}_prof_plain3.log 2>&1 && $NCU -k regex:"k_logit_psi_draw|k_gram_partial|k_gram_reduce|k_beta_draw" -s 12 -c 4 -o gpurun_out/prof_${T}_gibbs_p64 $B3 > gpurun_out/${T}_prof_ncu3.log 2>&1; echo "gibbs rc=$?"
B4="python tools/bench_models.py --nb-iters 2 --nb-N 2000000 --mlogit-iters 1"
$B4 > gpurun_out/${T}_prof_plain4.log 2>&1 && $NCU -k regex:"k_gram_partial<32|k_gram_partialILi32" -s 2 -c 1 -o gpurun_out/prof_${T}_gram_p256 $B4 > gpurun_out/${T}_prof_ncu4.log 2>&1; echo "p256 rc=$?"
$B1 > gpurun_out/${T}_prof_plain5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $B1 > gpurun_out/${T}_prof_ncu5.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/prof_${T}_*.ncu-rep
