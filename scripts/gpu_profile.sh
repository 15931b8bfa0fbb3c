# ncu evidence for the bench command: launch list + one full capture per hot kernel.
# usage (on the GPU box): bash scripts/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bma_reduce_cls -s 3 -c 1 -f -o gpurun_out/prof_reduce_$TAG $CMD > gpurun_out/ncu_reduce_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:row_sort -s 4 -c 1 -f -o gpurun_out/prof_sort_$TAG $CMD > gpurun_out/ncu_sort_$TAG.log 2>&1
ls -la gpurun_out/
tail -3 gpurun_out/plain_$TAG.log
