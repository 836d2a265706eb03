# The evidence of a round on ONE B200: GPU parity suite, smoke, sanitizers, reference arm, bench, profiles.
# Usage: bash scripts/gpu_round2.sh <tag>
TAG=${1:-r2}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
bash scripts/gpu_sanitize.sh ${TAG} 2>&1 | grep -E "exit|SUMMARY|passed|failed" 
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
tail -c 300 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
bash scripts/gpu_profile.sh ${TAG} 2>&1 | tail -8
