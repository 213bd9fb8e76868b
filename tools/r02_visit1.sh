#!/bin/bash
# Round 2, visit 1: evidence gaps of round 1 -- A/B of the opt-in switches, the gated (never-run) tests, compute-sanitizer.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().splitlines()[-1]); k=d['kernels']; print(sys.argv[1], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], {n: round(v['ms_per_step']*1e3,1) for n,v in k.items()})" $1 2>&1 | tail -1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
P="python bench.py --workload polyvore --steps 10 --warmup 3 --no-cpu --no-scoring"
M="python bench.py --workload modanet --no-cpu --no-scoring --no-fp32"
E="python bench.py --steps 1000 --warmup 50 --no-cpu --no-fp32 --no-scoring"
echo "== polyvore default"; timeout -s KILL 150 $P > gpurun_out/nv_poly.json 2> gpurun_out/nv_poly.err; pick gpurun_out/nv_poly.json
echo "== polyvore, persistent bulk store"; CODAE_TMA_STORE_PERSISTENT=1 timeout -s KILL 150 $P > gpurun_out/nv_poly_tma.json 2> gpurun_out/nv_poly_tma.err; pick gpurun_out/nv_poly_tma.json
echo "== modanet default"; timeout -s KILL 90 $M > gpurun_out/nv_modanet.json 2> gpurun_out/nv_modanet.err; pick gpurun_out/nv_modanet.json
echo "== modanet, layer-wise Adam"; CODAE_LAYERWISE_ADAM=1 timeout -s KILL 90 $M > gpurun_out/nv_modanet_lw.json 2> gpurun_out/nv_modanet_lw.err; pick gpurun_out/nv_modanet_lw.json
echo "== embedding default"; timeout -s KILL 90 $E > gpurun_out/nv_emb.json 2> gpurun_out/nv_emb.err; pick gpurun_out/nv_emb.json
echo "== embedding --deferred-update"; timeout -s KILL 90 $E --deferred-update > gpurun_out/nv_emb_deferred.json 2> gpurun_out/nv_emb_deferred.err; echo "rc=$?"; pick gpurun_out/nv_emb_deferred.json
echo "== embedding --no-pdl"; timeout -s KILL 90 $E --no-pdl > gpurun_out/nv_emb_nopdl.json 2> gpurun_out/nv_emb_nopdl.err; pick gpurun_out/nv_emb_nopdl.json
echo "== gated tests minus the chain kernel"
CODAE_EXPERIMENTAL=1 timeout -s KILL 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu -p no:cacheprovider -k "not chain" > gpurun_out/nv_experimental.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/nv_experimental.log
echo "== step timeline"; timeout -s KILL 120 python tools/step_timeline.py > gpurun_out/nv_timeline.txt 2>&1; echo "rc=$?"; tail -45 gpurun_out/nv_timeline.txt
echo "== sanitizer"
timeout -s KILL 1500 bash tools/gpu_sanitize.sh
echo "== chain kernel (last: protocol bugs must not cost the rest of the visit)"
timeout -s KILL 150 python tests/gpu_probe_chain.py > gpurun_out/nv_chain_probe.log 2>&1; echo "probe rc=$?"; tail -12 gpurun_out/nv_chain_probe.log
nvidia-smi --query-gpu=name,clocks.sm --format=csv
if grep -q "CHAIN OK" gpurun_out/nv_chain_probe.log; then
  CODAE_EXPERIMENTAL=1 timeout -s KILL 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu -p no:cacheprovider -k "chain" > gpurun_out/nv_experimental_chain.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/nv_experimental_chain.log
  echo "== embedding --chain"; timeout -s KILL 90 $E --chain > gpurun_out/nv_emb_chain.json 2> gpurun_out/nv_emb_chain.err; echo "rc=$?"; pick gpurun_out/nv_emb_chain.json
fi
echo "== done"
