#!/bin/bash
# Short GPU visit: probe + parity tests + default bench (no ncu).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== probe"; timeout -s KILL 150 python tests/gpu_probe_gemm.py > gpurun_out/probe.log 2>&1; prc=$?; echo "probe rc=$prc"; grep -c "bad=0/" gpurun_out/probe.log; grep -v "bad=0/" gpurun_out/probe.log | tail -12
if [ $prc -eq 137 ]; then echo "probe hung: stopping this visit"; exit 3; fi
run_pytest() { echo "== pytest $1"; timeout -s KILL 600 python -m pytest $2 -m gpu -q -p no:cacheprovider > gpurun_out/pytest_$1.log 2>&1; echo "rc=$?"; tail -${3:-15} gpurun_out/pytest_$1.log; }
run_pytest kernels tests/test_gpu_kernels.py 30
run_pytest training tests/test_gpu_training.py 30
run_pytest inference tests/test_gpu_inference.py 8
echo "== bench default"; timeout -s KILL 600 python bench.py --no-cpu > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -3 gpurun_out/bench_default.err
echo "== bench default, split-K off"; timeout -s KILL 600 python bench.py --no-cpu --no-scoring --no-fp32 --no-splitk > gpurun_out/bench_nosplit.json 2> gpurun_out/bench_nosplit.err; echo "rc=$?"; tail -3 gpurun_out/bench_nosplit.err
echo "== bench polyvore"; timeout -s KILL 600 python bench.py --workload polyvore --steps 10 --warmup 3 --no-cpu --no-scoring > gpurun_out/bench_polyvore.json 2> gpurun_out/bench_polyvore.err; echo "rc=$?"; tail -3 gpurun_out/bench_polyvore.err
echo "== bench polyvore, one tile per CTA"; timeout -s KILL 600 python bench.py --workload polyvore --steps 10 --warmup 3 --no-cpu --no-scoring --no-persistent > gpurun_out/bench_polyvore_np.json 2> gpurun_out/bench_polyvore_np.err; echo "rc=$?"; tail -3 gpurun_out/bench_polyvore_np.err
python - <<'PY'
import json
for f in ["bench_default.json","bench_nosplit.json","bench_polyvore.json","bench_polyvore_np.json"]:
    try: d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
    except Exception as e: print(f,"ERR",e); continue
    print("==",f,"value %.0f ms/step %.4f e2e %.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), "roofline", d["roofline"]["kernel"], "%.3f"%d["roofline"]["frac"], d["roofline"]["bound"])
    for k,v in d["kernels"].items(): print("     %-18s %8.3f ms/step %3d launches %8.2f us/launch %s"%(k,v["ms_per_step"],v["launches_per_step"],v["us_per_launch"], v.get("ms_per_step_back_to_back","")))
    if d.get("fp32_engine"): print("   fp32:", d["fp32_engine"])
    if d.get("scoring"): print("   scoring: %.3g scores/s frac %.3f"%(d["scoring"]["value"], d["scoring"]["roofline"]["frac"]))
PY
echo "== done"
