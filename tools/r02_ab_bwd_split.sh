#!/bin/bash
# A/B: how the concurrent dgrad (main stream) and wgrad (side stream) of the small-batch step share the 148 one-CTA-per-SM slots
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {  # workload dtype env...
  wl=$1; dt=$2; shift 2
  env "$@" timeout -s KILL 120 python bench.py --workload $wl --steps 1000 --warmup 50 --no-cpu --no-scoring --dtype $dt > gpurun_out/ab_t.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/ab_t.json')); print('$wl $dt $*', round(d['ms_per_step'],4), {n: round(v['ms_per_step']*1e3,1) for n,v in d['kernels'].items() if 'linear' in n})"
}
run embedding bf16 CODAE_DGRAD_BN=128
run embedding bf16 CODAE_DGRAD_BN=128 CODAE_DGRAD_MAXSPLIT=6
run embedding bf16 CODAE_DGRAD_BN=128 CODAE_DGRAD_MAXSPLIT=4
run embedding bf16 CODAE_DGRAD_BN=128 CODAE_WGRAD_BN=256
run embedding bf16 CODAE_DGRAD_BN=128 CODAE_DGRAD_MAXSPLIT=6 CODAE_WGRAD_BN=256
run embedding bf16 CODAE_DGRAD_BN=64 CODAE_DGRAD_MAXSPLIT=3
run embedding bf16 CODAE_DGRAD_BN=64 CODAE_DGRAD_MAXSPLIT=4
run embedding fp32 CODAE_DGRAD_BN=128 CODAE_DGRAD_MAXSPLIT=6
run embedding fp32 CODAE_DGRAD_BN=128 CODAE_DGRAD_MAXSPLIT=4
run modanet bf16 CODAE_DGRAD_BN=128 CODAE_DGRAD_MAXSPLIT=6
