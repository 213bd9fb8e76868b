#!/bin/bash
# Round 2, visit 4: scoring sweep with R rows per warp pass + staged merge, device sampler, full suite, default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
echo "== inference + training tests"
timeout -s KILL 600 python -m pytest tests/test_gpu_inference.py tests/test_gpu_training.py -q -m gpu -p no:cacheprovider > gpurun_out/v4_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/v4_tests.log
echo "== full gpu suite"
timeout -s KILL 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_inference.py --deselect tests/test_gpu_training.py > gpurun_out/v4_suite.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/v4_suite.log
echo "== default bench"
timeout -s KILL 900 python bench.py > gpurun_out/v4_bench_default.json 2> gpurun_out/v4_bench_default.err; echo "rc=$?"; tail -3 gpurun_out/v4_bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/v4_bench_default.json").read().splitlines()[-1])
print("primary", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "roofline", round(d["roofline"]["frac"], 3))
for k, v in d.get("secondary", {}).items():
    print("  ", k, v.get("ms_per_step"), v.get("engine", "")[:30], "e2e", v.get("e2e", {}).get("value"), v.get("error"))
s = d.get("scoring", {})
print("scoring f32", s.get("ms_per_sweep"), s.get("kernel_ms"), round(s["roofline"]["frac"], 3), "| bf16", s["bf16_catalog"]["ms_per_sweep"], s["bf16_catalog"]["kernel_ms"], round(s["bf16_catalog"]["roofline"]["frac"], 3))
print("cpu", d.get("cpu_baseline"))
PY
echo "== done"
