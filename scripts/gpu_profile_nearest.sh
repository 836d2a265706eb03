set -x
TAG=r1v8
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nearest_tiles_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_nearest_b256 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
ls -la gpurun_out | grep r1v8
