#!/bin/bash
# One short GPU-box visit for the opt-in paths (TMA bulk-store epilogue, norm-free clipped step): their parity tests, the
# WHOLE GPU suite with the bulk-store epilogue forced on, then the default bench command with each switch (A/B).
# Every stage has its own timeout; everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B="python bench.py --steps 1000 --warmup 50 --no-cpu --no-fp32 --no-scoring"
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().splitlines()[-1]); k=d['kernels']; print(sys.argv[1], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], {n: round(v['ms_per_step']*1e3,1) for n,v in k.items()})" $1 2>&1 | tail -1; }
echo "== new tests"; timeout -s KILL 120 python -m pytest tests/test_gpu_wgrad_sqnorm.py -q -m gpu -p no:cacheprovider > gpurun_out/exp_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/exp_tests.log
echo "== whole GPU suite, bulk-store epilogue on"; CODAE_TMA_STORE=1 timeout -s KILL 170 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/exp_suite_tma.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/exp_suite_tma.log
echo "== bench tma store"; timeout -s KILL 60 $B --tma-store > gpurun_out/exp_tma.json 2> gpurun_out/exp_tma.err; echo "rc=$?"; pick gpurun_out/exp_tma.json
echo "== bench tma store + sqnorm, Adam without PDL"; CODAE_ADAM_PARTIALS_PDL=0 timeout -s KILL 60 $B --wgrad-sqnorm --tma-store > gpurun_out/exp_sq_tma_nopdl.json 2> gpurun_out/exp_sq_tma_nopdl.err; echo "rc=$?"; pick gpurun_out/exp_sq_tma_nopdl.json
echo "== bench modanet tma store"; timeout -s KILL 60 python bench.py --workload modanet --no-cpu --no-scoring --no-fp32 --tma-store > gpurun_out/exp_modanet_tma.json 2> gpurun_out/exp_modanet_tma.err; echo "rc=$?"; pick gpurun_out/exp_modanet_tma.json
echo "== done"
