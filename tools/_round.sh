# One GPU call at the end of a change: parity tests, smoke, bench (both arms), stage times of the logit sweep.
#   gpurun --timeout 900 -- 'bash tools/_round.sh r17'          (NCU=1 adds the ncu launch list of bench.py)
mkdir -p gpurun_out
T=${1:-rXX}
timeout 480 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/${T}_pytest.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 150 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 330 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$? stdout lines: $(wc -l < gpurun_out/${T}_bench.json)"; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("value %.4g e2e %.4g bound %.4g clocks %s" % (d["value"], d["e2e"]["value"], d["e2e"]["pcie_bound_draws_per_s"], d["clocks"]))
PY
BL_GIBBS_TIMING=1 timeout 120 python tools/bench_gibbs.py --iters 100 > gpurun_out/${T}_gibbs.log 2>&1; tail -2 gpurun_out/${T}_gibbs.log
if [ -n "$NCU" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --no-extras --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
fi
