#!/bin/bash
# A/B of the persistent GEMM's raster group size (polyvore-shaped step).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in "16 0" "32 0" "64 0" "8 0"; do
  set -- $cfg
  CODAE_GROUP_M=$1 timeout -s KILL 200 python bench.py --workload polyvore --steps 10 --warmup 3 --no-cpu --no-scoring > gpurun_out/raster_$1_$2.json 2> gpurun_out/raster_$1_$2.err; echo "group_m=$1 l2_hint=$2 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/raster_$1_$2.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print("   ms/step %.4f  fwd %.1f  dgrad %.1f  wgrad %.1f us/launch"%(d["ms_per_step"],k["linear_fwd"]["us_per_launch"],k["linear_dgrad"]["us_per_launch"],k["linear_wgrad"]["us_per_launch"]))
except Exception as e: print("parse error", e)
PY
done
