#!/bin/bash
# Round 2, second profiling visit: the small-batch kernels after the geometry changes of visits 6-7 (128-wide dgrad tiles, multi-accumulator
# x3 epilogue) + the launch list of the default bench command with the final code.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
B="python bench.py --no-cpu --no-scoring --no-secondary --no-graph --steps 3 --warmup 3"
echo "== launch list of the default bench command"
python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/p2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file gpurun_out/p2_launches.csv python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/p2_launches.log 2>&1
echo "rc=$?"
echo "== embedding bf16 GEMMs (fwd, fwd, wgrad, dgrad, wgrad)"
$B --workload embedding --dtype bf16 > gpurun_out/p2_emb_plain.log 2>&1 &&
$NCU -k regex:tc05_gemm_kernel -s 95 -c 5 -o gpurun_out/prof_r2b_embedding_bf16_gemm -f $B --workload embedding --dtype bf16 > gpurun_out/p2_emb_gemm.log 2>&1
echo "rc=$?"
echo "== embedding fp32 (x3) GEMMs"
$B --workload embedding --dtype fp32 > gpurun_out/p2_x3_plain.log 2>&1 &&
$NCU -k regex:tc05_gemm_kernel -s 95 -c 5 -o gpurun_out/prof_r2b_embedding_fp32x3_gemm -f $B --workload embedding --dtype fp32 > gpurun_out/p2_x3_gemm.log 2>&1
echo "rc=$?"
echo "== launch list of one embedding.yaml fp32 step (all kernels of the step)"
ncu --metrics gpu__time_duration.sum --clock-control none -s 132 -c 66 --csv --log-file gpurun_out/p2_launches_x3.csv $B --workload embedding --dtype fp32 > gpurun_out/p2_launches_x3.log 2>&1
echo "rc=$?"
ls -la gpurun_out/prof_r2b_*.ncu-rep
echo "== done"
