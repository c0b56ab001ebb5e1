K='regex:k_xtv_stream|k_mlogit_offsets|k_beta_draw'
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 40 -c 3 -f -o gpurun_out/prof_r1_15_mlogit python tools/bench_models.py --nb-iters 0 --mlogit-iters 3 > gpurun_out/ncu15a.log 2>&1; echo rc=$?
