#!/bin/bash
# Visit for the K1 (corruption / loss) kernels: parity tests, then the default and the polyvore bench with per-kernel rooflines.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run_pytest() { echo "== pytest $1"; timeout -s KILL 500 python -m pytest $2 -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_$1.log 2>&1; echo "rc=$?"; tail -${3:-6} gpurun_out/pytest_$1.log; }
run_pytest kernels tests/test_gpu_kernels.py 8
run_pytest training tests/test_gpu_training.py 8
timeout -s KILL 300 python bench.py --no-cpu --no-scoring --no-fp32 > gpurun_out/bench_k1_default.json 2> gpurun_out/bench_k1_default.err; echo "bench default rc=$?"; tail -2 gpurun_out/bench_k1_default.err
timeout -s KILL 400 python bench.py --workload polyvore --steps 10 --warmup 3 --no-cpu --no-scoring > gpurun_out/bench_k1_polyvore.json 2> gpurun_out/bench_k1_polyvore.err; echo "bench polyvore rc=$?"; tail -2 gpurun_out/bench_k1_polyvore.err
timeout -s KILL 300 python bench.py --workload modanet --no-cpu --no-scoring --no-fp32 > gpurun_out/bench_k1_modanet.json 2> gpurun_out/bench_k1_modanet.err; echo "bench modanet rc=$?"; tail -2 gpurun_out/bench_k1_modanet.err
python - <<'PY'
import json
for f in ["bench_k1_default.json","bench_k1_modanet.json","bench_k1_polyvore.json"]:
    try: d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
    except Exception as e: print(f,"ERR",e); continue
    print("==",f,"value %.0f ms/step %.4f e2e %.0f"%(d["value"],d["ms_per_step"],d["e2e"]["value"]))
    for k,v in d["kernels"].items():
        r=d["rooflines"].get(k,{})
        print("     %-18s %8.3f ms/step %3d launches %8.2f us/launch  frac %.3f %s"%(k,v["ms_per_step"],v["launches_per_step"],v["us_per_launch"], r.get("frac",0), r.get("bound","")))
PY
