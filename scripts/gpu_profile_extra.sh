# ncu evidence for the kernels outside the bench step: diagnostics (KSD, ESS), conformal Bayes, narrow-head reductions,
# the GEMM with the row-statistics epilogue.  usage (on the GPU box): bash scripts/gpu_profile_extra.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
D="python scripts/bench_diagnostic.py"
timeout 300 $D > gpurun_out/plain_diag_$TAG.log 2>&1 || { echo "plain diagnostics run failed"; tail -5 gpurun_out/plain_diag_$TAG.log; }
# the c5 launches are the last ones of each kernel in the script: skip the warm-ups and smaller configs (7 launches per config)
timeout 600 $NCU -k regex:ksd_pair_sums -s 21 -o gpurun_out/prof_ksd_$TAG $D > gpurun_out/ncu_ksd_$TAG.log 2>&1
timeout 600 $NCU -k regex:ess_kernel -s 21 -o gpurun_out/prof_ess_$TAG $D > gpurun_out/ncu_ess_$TAG.log 2>&1
S="python scripts/bench_scoring.py --iters 1"
timeout 300 $S > gpurun_out/plain_scoring_$TAG.log 2>&1
timeout 600 $NCU -k regex:cb_count -s 2 -o gpurun_out/prof_cb_$TAG $S > gpurun_out/ncu_cb_$TAG.log 2>&1
C="python scripts/bench_configs.py"
timeout 300 $C > gpurun_out/plain_configs_$TAG.log 2>&1
timeout 600 $NCU -k regex:bma_reduce_cls_narrow -s 30 -o gpurun_out/prof_narrow_$TAG $C > gpurun_out/ncu_narrow_$TAG.log 2>&1
G="python scripts/bench_last_layer.py --ex --no-cublas --iters 1"
timeout 300 $G > gpurun_out/plain_tcex_$TAG.log 2>&1
timeout 600 $NCU -k regex:last_layer_tc_kernel -s 7 -o gpurun_out/prof_tcex_$TAG $G > gpurun_out/ncu_tcex_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep
for f in gpurun_out/plain_diag_$TAG.log gpurun_out/plain_scoring_$TAG.log gpurun_out/plain_tcex_$TAG.log; do tail -n 1 $f | cut -c1-400; done
