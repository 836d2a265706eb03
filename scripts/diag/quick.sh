# quick GPU round: the contact / objective tests, then the bench's own step with the per-kernel times
python -m pytest tests/test_contact_gpu.py tests/test_objective_gpu.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']; print(round(d['ms_per_step'],4), {a:round(b,4) for a,b in k.items()})"
python scripts/diag/small_batch.py 32 2>&1 | grep "B="
