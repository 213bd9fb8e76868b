#!/bin/bash
# Short GPU visit for the swap-reconstruction scorer: parity tests + default bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest inference"; timeout -s KILL 420 python -m pytest tests/test_gpu_inference.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_inference.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_inference.log
echo "== bench default"; timeout -s KILL 420 python bench.py --no-cpu --no-fp32 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -5 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.0f ms/step %.4f e2e %.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]))
print(json.dumps(d["scoring"], indent=1))
PY
