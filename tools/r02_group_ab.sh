#!/bin/bash
# Raster group of the CTA-pair kernel: step time (A/B, two runs each) and DRAM traffic (ncu --set full, five launches of one step)
# with CODAE_GROUP_M=32 (16 pair-rows per group) against the default 16 (8 pair-rows).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-scoring --no-secondary --steps 20 --warmup 5"
for rnd in 0 1; do for g in 16 32; do
  CODAE_GROUP_M=$g timeout 120 $B > gpurun_out/group_${g}_${rnd}.json 2> gpurun_out/group_${g}_${rnd}.err; echo "group_m=$g round $rnd rc=$?"
  python -c "import json;d=json.loads(open('gpurun_out/group_${g}_${rnd}.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['clocks'],{k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if 'linear' in k})"
done; done
NB="python bench.py --no-cpu --no-scoring --no-secondary --no-graph --steps 3 --warmup 3"
CODAE_GROUP_M=32 timeout 240 ncu --set full --clock-control none --import-source on -k regex:tc05_gemm_persistent -s 95 -c 5 -o gpurun_out/prof_r2d_polyvore_gemm_pair_group32 -f $NB --workload polyvore > gpurun_out/group32_ncu.log 2>&1
echo "ncu rc=$?"
