#!/bin/bash
# The A/B of the round-1 switches, re-runnable with the current flag names (profiles/r01_ab_*.json were taken with the first
# version of this script, when the switches were still opt-in): TMA bulk-store epilogue, norm-free clipped step, optimizer
# launch with / without programmatic dependent launch.  Every stage has its own timeout; everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B="python bench.py --steps 1000 --warmup 50 --no-cpu --no-fp32 --no-scoring"
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().splitlines()[-1]); k=d['kernels']; print(sys.argv[1], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], {n: round(v['ms_per_step']*1e3,1) for n,v in k.items()})" $1 2>&1 | tail -1; }
echo "== default (bulk store, norm-free update, optimizer without PDL)"; timeout -s KILL 90 $B > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err; echo "rc=$?"; pick gpurun_out/ab_default.json
echo "== --no-tma-store"; timeout -s KILL 90 $B --no-tma-store > gpurun_out/ab_no_tma.json 2> gpurun_out/ab_no_tma.err; echo "rc=$?"; pick gpurun_out/ab_no_tma.json
echo "== --no-wgrad-sqnorm (cooperative norm + Adam)"; timeout -s KILL 90 $B --no-wgrad-sqnorm > gpurun_out/ab_no_sq.json 2> gpurun_out/ab_no_sq.err; echo "rc=$?"; pick gpurun_out/ab_no_sq.json
echo "== optimizer launched as a programmatic dependent"; CODAE_ADAM_PDL=1 timeout -s KILL 90 $B > gpurun_out/ab_adam_pdl.json 2> gpurun_out/ab_adam_pdl.err; echo "rc=$?"; pick gpurun_out/ab_adam_pdl.json
echo "== modanet default / --no-tma-store"
M="python bench.py --workload modanet --no-cpu --no-scoring --no-fp32"
timeout -s KILL 90 $M > gpurun_out/ab_modanet.json 2> gpurun_out/ab_modanet.err; pick gpurun_out/ab_modanet.json
timeout -s KILL 90 $M --no-tma-store > gpurun_out/ab_modanet_no_tma.json 2> gpurun_out/ab_modanet_no_tma.err; pick gpurun_out/ab_modanet_no_tma.json
echo "== done"
