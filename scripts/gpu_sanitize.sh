# compute-sanitizer memcheck + racecheck over the small-size GPU parity tests and smoke().
# Usage: bash scripts/gpu_sanitize.sh <tag>     (summaries land in gpurun_out/<tag>_{memcheck,racecheck}*.log)
TAG=${1:-r2}
SEL='not full and not batch64 and not touching and not reproducible and not reference_formulation and not large_batch'
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${TAG}_memcheck_tests.log \
    python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/${TAG}_memcheck_tests.out 2>&1
echo "memcheck tests exit $?"; tail -3 gpurun_out/${TAG}_memcheck_tests.out; tail -3 gpurun_out/${TAG}_memcheck_tests.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${TAG}_memcheck_smoke.log \
    python __graft_entry__.py smoke > gpurun_out/${TAG}_memcheck_smoke.out 2>&1
echo "memcheck smoke exit $?"; tail -2 gpurun_out/${TAG}_memcheck_smoke.out; tail -2 gpurun_out/${TAG}_memcheck_smoke.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/${TAG}_racecheck_tests.log \
    python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/${TAG}_racecheck_tests.out 2>&1
echo "racecheck tests exit $?"; tail -3 gpurun_out/${TAG}_racecheck_tests.out; tail -3 gpurun_out/${TAG}_racecheck_tests.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/${TAG}_racecheck_smoke.log \
    python __graft_entry__.py smoke > gpurun_out/${TAG}_racecheck_smoke.out 2>&1
echo "racecheck smoke exit $?"; tail -2 gpurun_out/${TAG}_racecheck_smoke.out; tail -2 gpurun_out/${TAG}_racecheck_smoke.log
