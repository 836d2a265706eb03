# compute-sanitizer over the tests that drive the nearest-vertex kernels (tile kernel, mixed kernel, single-query
# kernel, tile-pair summary) and the hierarchy pack, plus smoke()
SEL='within or nearest or shifted or pruned or fused_iteration'
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/r2v10_${tool}_nearest.log \
      python -m pytest tests/test_contact_gpu.py tests/test_objective_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/r2v10_${tool}_nearest.out 2>&1
  echo "$tool exit $?"; tail -2 gpurun_out/r2v10_${tool}_nearest.out; tail -2 gpurun_out/r2v10_${tool}_nearest.log
done
