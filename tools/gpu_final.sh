#!/bin/bash
# Last short visit of a round: the driver's GPU test command on the final defaults, then the default bench line and smoke().
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest (driver command)"; timeout -s KILL 100 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/final_pytest_gpu.log
echo "== bench default (no CPU leg)"; timeout -s KILL 70 python bench.py --no-cpu > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "rc=$?"; tail -2 gpurun_out/final_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/final_bench_default.json').read().splitlines()[-1])
print('ms/step %.4f value %.0f e2e %.0f launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches']))
print({n: round(v['ms_per_step']*1e3,1) for n,v in d['kernels'].items()}); print(d['roofline']); print(d['scoring']['value'] if d['scoring'] else None)"
echo "== smoke"; timeout -s KILL 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== done"
