#!/bin/bash
# Round 2, third profiling visit: the CTA-pair persistent kernel (cta_group::2) -- `ncu --set full` of five launches of one polyvore
# step (fwd, fwd, wgrad, dgrad, wgrad) + the launch list of the default bench command with the final code.  ONE gpurun call.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
B="python bench.py --no-cpu --no-scoring --no-secondary --no-graph --steps 3 --warmup 3"
echo "== polyvore: CTA-pair persistent GEMM (fwd, fwd, wgrad, dgrad, wgrad of one step)"
$B --workload polyvore > gpurun_out/p3_poly_plain.log 2>&1 &&
timeout 240 $NCU -k regex:tc05_gemm_persistent -s 95 -c 5 -o gpurun_out/prof_r2c_polyvore_gemm_pair -f $B --workload polyvore > gpurun_out/p3_poly_gemm.log 2>&1
echo "rc=$?"
echo "== launch list of the default bench command"
python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/p3_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file gpurun_out/p3_launches.csv python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/p3_launches.log 2>&1
echo "rc=$?"
ls -la gpurun_out/prof_r2c_*.ncu-rep
tail -2 gpurun_out/p3_plain.log | cut -c1-600
echo "== done"
