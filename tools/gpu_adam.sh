#!/bin/bash
# A/B: plain Adam (no clipping configured) through adam_kernel vs through the cooperative norm+update kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in 0 1; do
CODAE_ADAM_VIA_COOP=$v timeout -s KILL 200 python bench.py --workload modanet --no-cpu --no-scoring --no-fp32 > gpurun_out/adam_$v.json 2> gpurun_out/adam_$v.err; echo "coop=$v rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/adam_$v.json").read().strip().splitlines()[-1])
print("  ms/step %.4f"%d["ms_per_step"], {k:round(x["us_per_launch"],1) for k,x in d["kernels"].items() if "adam" in k})
PY
done
