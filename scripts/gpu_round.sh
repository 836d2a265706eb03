# One GPU-box visit: tests, smoke, bench (ours + reference arm).  Usage: bash scripts/gpu_round.sh [tests|notests]
set -x
if [ "${1:-tests}" = "tests" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
