# Runs every GPU test file in its own process (a hang in one cannot hide the others), then smoke + bench.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in prng sampler scoring predictive last_layer_tc golden mirror_api prob_model regression diagnostic guards output_calib maxcov calib_model multivalid full_size ${EXTRA_TESTS}; do
  [ -f tests/test_gpu_$f.py ] || continue
  timeout -k 10 ${TEST_TIMEOUT:-600} python -m pytest tests/test_gpu_$f.py -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "test_$f exit $?" >> gpurun_out/summary.txt
done
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout -k 10 900 python bench.py ${BENCH_ARGS:---steps 20 --warmup 3} > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/test_*.log; do echo "== $f"; tail -n 4 $f; done
tail -n 3 gpurun_out/bench.log
