# ncu evidence for the bench command: launch list + one full capture per hot kernel (reduce, sort, GEMM, sampler).
# usage (on the GPU box): bash scripts/gpu_profile.sh <tag>      every command runs under its own timeout
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras"
timeout 200 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bma_reduce_cls -s 3 -c 1 -f -o gpurun_out/prof_reduce_$TAG $CMD > gpurun_out/ncu_reduce_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:aps_sets_select -s 4 -c 1 -f -o gpurun_out/prof_select_$TAG $CMD > gpurun_out/ncu_select_$TAG.log 2>&1
# the full sort over every row (what the select kernel replaces where it can decide): validation rows from ONE sample
timeout 300 ncu --set full --clock-control none --import-source on -k regex:row_sort32_warp_kernel -s 3 -c 1 -f -o gpurun_out/prof_sort_$TAG $CMD --val-samples 1 > gpurun_out/ncu_sort_$TAG.log 2>&1
GEMM="python scripts/bench_last_layer.py --D 2048 --out f32 --iters 1 --no-cublas"
timeout 200 $GEMM > gpurun_out/plain_tc_$TAG.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:last_layer_tc -s 1 -c 1 -f -o gpurun_out/prof_tc_$TAG $GEMM > gpurun_out/ncu_tc_$TAG.log 2>&1
SAMP="python scripts/bench_sampler.py --steps 3"
timeout 200 $SAMP > gpurun_out/plain_sampler_$TAG.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgmcmc_step_kernel -s 3 -c 1 -f -o gpurun_out/prof_sampler_$TAG $SAMP > gpurun_out/ncu_sampler_$TAG.log 2>&1
ls -la gpurun_out/ | tail -20
for f in gpurun_out/plain_$TAG.log gpurun_out/plain_tc_$TAG.log gpurun_out/plain_sampler_$TAG.log; do tail -n 2 $f; done
