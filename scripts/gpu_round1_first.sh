mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in prng sampler scoring predictive; do
  timeout -k 10 600 python -m pytest tests/test_gpu_$f.py -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "test_$f exit $?" >> gpurun_out/summary.txt
done
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/test_*.log
tail -3 gpurun_out/bench.log
