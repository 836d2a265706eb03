python -m pytest tests/test_contact_gpu.py -x -q -m gpu 2>&1 | tail -3
q() { python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],4), 'pack', k['cluster_pack_kernel'], 'nn', k['nearest_kernel'], 'wind', k['winding_kernel'], 'bwd', k['lbs_backward_kernels'])"; }
q NOCAP
