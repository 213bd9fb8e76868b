#!/bin/bash
# Multi-GPU validation of exactly what the driver's scaling run launches: DP parity check, then the default bench line at N GPUs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
echo "== dist check ($N GPUs)"; CODAE_DP_TIMEOUT_S=20 timeout -s KILL 500 $TR 29511 tests/dist_gpu_check.py > gpurun_out/scale_check_${N}gpu.log 2>&1; echo "rc=$?"; grep -E "DP mode|sharded|DIST CHECK|Error|error|Traceback" gpurun_out/scale_check_${N}gpu.log | tail -24
echo "== default bench ($N GPUs)"
CODAE_DP_TIMEOUT_S=20 timeout -s KILL 600 $TR 29512 bench.py --gpus $N > gpurun_out/scale_bench_${N}gpu.json 2> gpurun_out/scale_bench_${N}gpu.err; echo "rc=$?"; tail -3 gpurun_out/scale_bench_${N}gpu.err | cut -c1-300
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open("gpurun_out/scale_bench_%sgpu.json" % n).read().splitlines()[-1])
print("primary", d["config"]["workload"][:40], "dp", d.get("dp_mode"), "ms/step %.4f" % d["ms_per_step"], "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"])
for k, v in d.get("secondary", {}).items():
    print("  ", k, v.get("ms_per_step"), v.get("dp_mode"), "value", v.get("value"), v.get("error"))
s = d.get("scoring", {})
print("scoring f32 ms", s.get("ms_per_sweep"), "kernel", s.get("kernel_ms"), "| bf16", s.get("bf16_catalog", {}).get("ms_per_sweep"), "| swaps", s.get("swap_reconstruction", {}).get("value"))
PY
echo "== done"
