# compute-sanitizer on the small workload (one tool per gpurun call, as the profiling guide asks):
#   gpurun --timeout 900 -- 'bash tools/_sanitize.sh memcheck r2'
TOOL=${1:-memcheck}; T=${2:-rX}
mkdir -p gpurun_out
export BL_SANITIZE_NUM=60000
python tools/sanitize_small.py > gpurun_out/${T}_sanitize_plain.log 2>&1 && \
timeout 800 compute-sanitizer --tool $TOOL --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/${T}_sanitize_${TOOL}.log 2>&1
echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|hybrid|devroye|gibbs|Error|RACECHECK SUMMARY" gpurun_out/${T}_sanitize_${TOOL}.log | tail -12
