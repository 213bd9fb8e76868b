#!/bin/bash
# One short GPU-box visit for the two opt-in paths (norm-free clipped step, TMA bulk-store epilogue): their parity tests,
# then the default bench command with each switch (A/B against the default), then the bf16 training tests with the
# norm-free step forced on.  Every stage has its own timeout; everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B="python bench.py --steps 1000 --warmup 50 --no-cpu --no-fp32 --no-scoring"
echo "== tests"; timeout -s KILL 200 python -m pytest tests/test_gpu_wgrad_sqnorm.py -q -m gpu -p no:cacheprovider > gpurun_out/exp_tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/exp_tests.log
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().splitlines()[-1]); k=d['kernels']; print(sys.argv[1], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], {n: round(v['ms_per_step']*1e3,1) for n,v in k.items()})" $1 2>&1 | tail -1; }
echo "== bench sqnorm"; timeout -s KILL 90 $B --wgrad-sqnorm > gpurun_out/exp_sq.json 2> gpurun_out/exp_sq.err; echo "rc=$?"; pick gpurun_out/exp_sq.json
echo "== bench default"; timeout -s KILL 90 $B > gpurun_out/exp_default.json 2> gpurun_out/exp_default.err; echo "rc=$?"; pick gpurun_out/exp_default.json
echo "== bench sqnorm + tma store"; timeout -s KILL 90 $B --wgrad-sqnorm --tma-store > gpurun_out/exp_sq_tma.json 2> gpurun_out/exp_sq_tma.err; echo "rc=$?"; pick gpurun_out/exp_sq_tma.json
echo "== bench tma store"; timeout -s KILL 90 $B --tma-store > gpurun_out/exp_tma.json 2> gpurun_out/exp_tma.err; echo "rc=$?"; pick gpurun_out/exp_tma.json
echo "== training tests, norm-free step forced on"; CODAE_WGRAD_SQNORM=1 timeout -s KILL 200 python -m pytest tests/test_gpu_training.py -q -m gpu -p no:cacheprovider -k "bf16 or full_size or fused_step_matches" > gpurun_out/exp_training.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/exp_training.log
echo "== polyvore sqnorm"; timeout -s KILL 120 python bench.py --workload polyvore --steps 10 --warmup 3 --no-cpu --no-scoring --wgrad-sqnorm > gpurun_out/exp_poly_sq.json 2> gpurun_out/exp_poly_sq.err; echo "rc=$?"; pick gpurun_out/exp_poly_sq.json
echo "== done"
