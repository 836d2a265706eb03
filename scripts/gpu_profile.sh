# One GPU-box visit for the profiles: launch list of a short bench run + ncu --set full captures of the
# dominant kernels inside the bench itself (B=256).  Usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r2}
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:winding_cluster_kernel -s 4 -c 1 -f \
    -o gpurun_out/${TAG}_winding_b256 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:nearest_tiles_kernel -s 4 -c 1 -f \
    -o gpurun_out/${TAG}_nearest_b256 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:lbs_skin_tc_kernel|lbs_tc_bwd_kernel|cluster_pack_top_kernel|segment_whitelist_kernel|lbs_bwd_vertex_kernel" -s 15 -c 5 -f \
    -o gpurun_out/${TAG}_lbs_pack_b256 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
ls -la gpurun_out/ | grep ${TAG}
