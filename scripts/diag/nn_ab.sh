set -x
python -m pytest tests/test_contact_gpu.py -x -q -m gpu -k "nearest or pruned or tiles" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CAP10', d['ms_per_step'], d['roofline']['nearest_kernel_avg_ms'], d['roofline']['avg_launch_ms'])"
sed -i 's/__launch_bounds__(NT_WARPS \* 32, 10)/__launch_bounds__(NT_WARPS * 32)/' tuch_b200/csrc/nearest_tiles.cu
python -m tuch_b200.build | tail -1
python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NOCAP', d['ms_per_step'], d['roofline']['nearest_kernel_avg_ms'], d['roofline']['avg_launch_ms'])"
python scripts/time_contact.py 2>&1 | tail -6
