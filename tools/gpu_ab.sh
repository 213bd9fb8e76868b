#!/bin/bash
# A/B visit: training parity tests, then the default bench with and without one switch ($1, e.g. --no-prefetch).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== probe"; timeout -s KILL 120 python tests/gpu_probe_gemm.py > gpurun_out/probe.log 2>&1; prc=$?; echo "probe rc=$prc"; grep -c "bad=0/" gpurun_out/probe.log; grep -v "bad=0/" gpurun_out/probe.log | tail -5
if [ $prc -eq 137 ]; then echo "probe hung: stopping this visit"; exit 3; fi
echo "== pytest training"; timeout -s KILL 400 python -m pytest tests/test_gpu_training.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_training.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_training.log
for tag in on off; do
  flag=""; [ $tag = off ] && flag="$1"
  timeout -s KILL 300 python bench.py --no-cpu --no-scoring --no-fp32 $flag > gpurun_out/bench_ab_$tag.json 2> gpurun_out/bench_ab_$tag.err; echo "bench $tag rc=$?"; tail -2 gpurun_out/bench_ab_$tag.err
done
timeout -s KILL 300 python bench.py --workload modanet --no-cpu --no-scoring --no-fp32 > gpurun_out/bench_ab_modanet.json 2> gpurun_out/bench_ab_modanet.err; echo "bench modanet rc=$?"
python - <<'PY'
import json
for f in ["bench_ab_on.json","bench_ab_off.json","bench_ab_modanet.json"]:
    try: d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
    except Exception as e: print(f,"ERR",e); continue
    print("==",f,"value %.0f ms/step %.4f e2e %.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]))
    for k,v in d["kernels"].items(): print("     %-18s %8.3f ms/step %3d launches %8.2f us/launch %s"%(k,v["ms_per_step"],v["launches_per_step"],v["us_per_launch"], v.get("ms_per_step_back_to_back","")))
PY
