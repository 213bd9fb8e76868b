#!/bin/bash
# Multi-GPU visit: DP parity (both schedules) + bench lines per schedule.  Usage: bash tools/r02_dist.sh N [quick]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
echo "== dist check ($N GPUs)"; CODAE_DP_TIMEOUT_S=20 timeout -s KILL 400 $TR 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1; echo "rc=$?"; grep -E "DP mode|sharded|DIST CHECK|Error|error|Traceback" gpurun_out/dist_check_${N}gpu.log | tail -20
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().splitlines()[-1]); print(sys.argv[1], 'dp_mode', d.get('dp_mode'), 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], {n: round(v['ms_per_step']*1e3,1) for n,v in d['kernels'].items()})" $1 2>&1 | tail -1; }
for MODE in peer nccl; do
  for WL in embedding polyvore; do
    echo "== bench $WL --dp-mode $MODE ($N GPUs)"
    CODAE_DP_TIMEOUT_S=20 timeout -s KILL 300 $TR 29512 bench.py --gpus $N --workload $WL --dp-mode $MODE --no-cpu --no-scoring --no-secondary > gpurun_out/dist_${WL}_${MODE}_${N}gpu.json 2> gpurun_out/dist_${WL}_${MODE}_${N}gpu.err; echo "rc=$?"; tail -2 gpurun_out/dist_${WL}_${MODE}_${N}gpu.err | cut -c1-300; pick gpurun_out/dist_${WL}_${MODE}_${N}gpu.json
  done
done
echo "== done"
