#!/bin/bash
# First GPU visit of the next round: A/B of the opt-in switches that are parity-checked but not yet timed, then the
# launch list + full ncu capture of the current default path (the committed ones predate the bulk-store epilogue).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().splitlines()[-1]); k=d['kernels']; print(sys.argv[1], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], {n: round(v['ms_per_step']*1e3,1) for n,v in k.items()})" $1 2>&1 | tail -1; }
P="python bench.py --workload polyvore --steps 10 --warmup 3 --no-cpu --no-scoring"
M="python bench.py --workload modanet --no-cpu --no-scoring --no-fp32"
echo "== polyvore default"; timeout -s KILL 120 $P > gpurun_out/nv_poly.json 2> gpurun_out/nv_poly.err; pick gpurun_out/nv_poly.json
echo "== polyvore, persistent bulk store"; CODAE_TMA_STORE_PERSISTENT=1 timeout -s KILL 120 $P > gpurun_out/nv_poly_tma.json 2> gpurun_out/nv_poly_tma.err; pick gpurun_out/nv_poly_tma.json
echo "== modanet default"; timeout -s KILL 90 $M > gpurun_out/nv_modanet.json 2> gpurun_out/nv_modanet.err; pick gpurun_out/nv_modanet.json
echo "== modanet, layer-wise Adam"; CODAE_LAYERWISE_ADAM=1 timeout -s KILL 90 $M > gpurun_out/nv_modanet_lw.json 2> gpurun_out/nv_modanet_lw.err; pick gpurun_out/nv_modanet_lw.json
echo "== experimental: chain kernel (never run on a GPU when it was written): probe, tests, A/B"
timeout -s KILL 150 python tests/gpu_probe_chain.py > gpurun_out/nv_chain_probe.log 2>&1; echo "probe rc=$?"; tail -12 gpurun_out/nv_chain_probe.log
CODAE_EXPERIMENTAL=1 timeout -s KILL 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu -p no:cacheprovider > gpurun_out/nv_experimental.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/nv_experimental.log
E="python bench.py --steps 1000 --warmup 50 --no-cpu --no-fp32 --no-scoring"
echo "== embedding default"; timeout -s KILL 90 $E > gpurun_out/nv_emb.json 2> gpurun_out/nv_emb.err; pick gpurun_out/nv_emb.json
echo "== embedding --chain"; timeout -s KILL 90 $E --chain > gpurun_out/nv_emb_chain.json 2> gpurun_out/nv_emb_chain.err; echo "rc=$?"; pick gpurun_out/nv_emb_chain.json
echo "== embedding --deferred-update (update of step s beside the forward pass of step s+1)"; timeout -s KILL 90 $E --deferred-update > gpurun_out/nv_emb_deferred.json 2> gpurun_out/nv_emb_deferred.err; echo "rc=$?"; pick gpurun_out/nv_emb_deferred.json
echo "== embedding --no-pdl"; timeout -s KILL 90 $E --no-pdl > gpurun_out/nv_emb_nopdl.json 2> gpurun_out/nv_emb_nopdl.err; pick gpurun_out/nv_emb_nopdl.json
echo "== step timeline (stamp kernels around every call, per stream)"; timeout -s KILL 120 python tools/step_timeline.py > gpurun_out/nv_timeline.txt 2>&1; echo "rc=$?"; tail -40 gpurun_out/nv_timeline.txt
echo "== ncu launch list (default command, short)"
CMD="python bench.py --steps 3 --warmup 3 --no-scoring --no-cpu --no-fp32 --no-graph"
timeout -s KILL 200 $CMD > gpurun_out/plain.log 2>&1 && \
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"tc05_gemm_kernel|adam_partials" -s 100 -c 10 -o gpurun_out/prof_gemm_small $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
echo "== done"
