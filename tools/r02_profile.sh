#!/bin/bash
# Round 2 profiling visit: launch list of the bench command + one `ncu --set full` capture per final kernel (ONE gpurun call).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
B="python bench.py --no-cpu --no-scoring --no-secondary --no-graph --steps 3 --warmup 3"
echo "== launch list of the default bench command (first 700 launches: polyvore steps)"
python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file gpurun_out/p_launches.csv python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/p_launches.log 2>&1
echo "rc=$?"
echo "== polyvore: persistent GEMM (fwd, fwd, wgrad, dgrad, wgrad of one step) + Adam"
$B --workload polyvore > gpurun_out/p_poly_plain.log 2>&1 &&
$NCU -k regex:tc05_gemm_persistent -s 95 -c 5 -o gpurun_out/prof_r2_polyvore_gemm -f $B --workload polyvore > gpurun_out/p_poly_gemm.log 2>&1
echo "rc=$?"
$NCU -k regex:adam_partials -s 3 -c 1 -o gpurun_out/prof_r2_polyvore_adam -f $B --workload polyvore > gpurun_out/p_poly_adam.log 2>&1
echo "rc=$?"
echo "== embedding bf16: split-K GEMM (fwd, wgrad, dgrad) + Adam"
$B --workload embedding --dtype bf16 > gpurun_out/p_emb_plain.log 2>&1 &&
$NCU -k regex:tc05_gemm_kernel -s 95 -c 5 -o gpurun_out/prof_r2_embedding_bf16_gemm -f $B --workload embedding --dtype bf16 > gpurun_out/p_emb_gemm.log 2>&1
echo "rc=$?"
$NCU -k regex:adam_partials -s 3 -c 1 -o gpurun_out/prof_r2_embedding_bf16_adam -f $B --workload embedding --dtype bf16 > gpurun_out/p_emb_adam.log 2>&1
echo "rc=$?"
echo "== embedding fp32 (three-plane engine)"
$B --workload embedding --dtype fp32 > gpurun_out/p_x3_plain.log 2>&1 &&
$NCU -k regex:tc05_gemm_kernel -s 95 -c 5 -o gpurun_out/prof_r2_embedding_fp32x3_gemm -f $B --workload embedding --dtype fp32 > gpurun_out/p_x3_gemm.log 2>&1
echo "rc=$?"
echo "== scoring (f32 + bf16 catalog)"
python tools/prof_scoring.py 4000000 > gpurun_out/p_score_plain.log 2>&1 &&
$NCU -k regex:score_topk -s 2 -c 1 -o gpurun_out/prof_r2_scoring_f32 -f python tools/prof_scoring.py 4000000 > gpurun_out/p_score_f32.log 2>&1
echo "rc=$?"
$NCU -k regex:score_topk -s 5 -c 1 -o gpurun_out/prof_r2_scoring_bf16 -f python tools/prof_scoring.py 4000000 > gpurun_out/p_score_bf16.log 2>&1
echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
echo "== done"
