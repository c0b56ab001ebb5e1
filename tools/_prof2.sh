# Profiling of the kernels changed in the second half of round 2 (one GPU): ncu --set full captures on the current tree.
#   gpurun --timeout 1200 -- 'bash tools/_prof2.sh r2b'
# Every ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
T=${1:-rX}
NCU="ncu --set full --clock-control none --import-source on -f"
B1="python tools/bench_gibbs.py --iters 6 --warmup 2 --two-pass"
$B1 > gpurun_out/${T}_prof_plain1.log 2>&1 && $NCU -k regex:"k_logit_psi_draw|k_gram_partial|k_gram_reduce|k_beta_draw" -s 12 -c 4 -o gpurun_out/prof_${T}_gibbs_p64 $B1 > gpurun_out/${T}_prof_ncu1.log 2>&1; echo "gibbs rc=$?"
B2="python tools/bench_gibbs.py --iters 6 --warmup 2 --constrained --N 200000"
$B2 > gpurun_out/${T}_prof_plain2.log 2>&1 && $NCU -k regex:"k_beta_draw|k_tn_normals" -s 8 -c 2 -o gpurun_out/prof_${T}_beta_constrained $B2 > gpurun_out/${T}_prof_ncu2.log 2>&1; echo "constrained rc=$?"
B3="python tools/bench_models.py --nb-iters 0 --mlogit-iters 2"
$B3 > gpurun_out/${T}_prof_plain3.log 2>&1 && $NCU -k regex:"k_xbeta_mma|k_devroye_refill|k_cls_|k_gram_partial" -s 30 -c 6 -o gpurun_out/prof_${T}_mlogit $B3 > gpurun_out/${T}_prof_ncu3.log 2>&1; echo "mlogit rc=$?"
B4="python tools/bench_gibbs.py --iters 6 --warmup 2 --N 125000"
$B4 > gpurun_out/${T}_prof_plain4.log 2>&1 && $NCU -k regex:"k_logit_sweep" -s 4 -c 1 -o gpurun_out/prof_${T}_k3_125k $B4 > gpurun_out/${T}_prof_ncu4.log 2>&1; echo "k3 rc=$?"
ls -la gpurun_out/prof_${T}_*.ncu-rep
# the reports of one call must fit gpurun's 64 MiB merge limit: condense them here, keep the small ones
for r in gibbs_p64 beta_constrained mlogit k3_125k; do
  python tools/ncu_summary.py gpurun_out/prof_${T}_$r.ncu-rep > gpurun_out/${T}_ncu_summary_$r.txt 2>/dev/null
done
rm -f gpurun_out/prof_${T}_gibbs_p64.ncu-rep gpurun_out/prof_${T}_mlogit.ncu-rep
