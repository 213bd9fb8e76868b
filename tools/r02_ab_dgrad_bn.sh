#!/bin/bash
# A/B: tile width of the input-gradient contraction at small batch (96 CTAs at BN = 128 leave SMs to the weight gradient on the side stream)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "embedding fp32 0" "embedding fp32 128" "embedding bf16 0" "embedding bf16 128" "modanet bf16 0" "modanet bf16 128"; do
  set -- $cfg
  CODAE_DGRAD_BN=$3 timeout -s KILL 120 python bench.py --workload $1 --steps 1000 --warmup 50 --no-cpu --no-scoring --dtype $2 > gpurun_out/ab_dgrad_$1_$2_$3.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/ab_dgrad_$1_$2_$3.json')); print('$1 $2 dgrad_bn=$3', round(d['ms_per_step'],4), {n: round(v['ms_per_step']*1e3,1) for n,v in d['kernels'].items()})"
done
