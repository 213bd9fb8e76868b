#!/bin/bash
# A/B of the NVLS (multimem) path of the data-parallel kernel.  Usage: bash tools/r02_nvls_ab.sh N
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
echo "== dist check ($N GPUs, NVLS on)"; CODAE_DP_TIMEOUT_S=20 timeout -s KILL 500 $TR 29511 tests/dist_gpu_check.py > gpurun_out/nvls_check_${N}gpu.log 2>&1; echo "rc=$?"; grep -E "DP |sharded|DIST CHECK|Error|error|Traceback" gpurun_out/nvls_check_${N}gpu.log | tail -24
for nv in 1 0; do
  for cfg in "embedding bf16" "embedding fp32" "polyvore bf16"; do
    set -- $cfg
    CODAE_DP_NVLS=$nv CODAE_DP_TIMEOUT_S=20 timeout -s KILL 300 $TR 29512 bench.py --gpus $N --workload $1 --dtype $2 --no-cpu --no-scoring --no-secondary > gpurun_out/nvls_${nv}_$1_$2_${N}gpu.json 2> gpurun_out/nvls_t.err
    python -c "
import json; d=json.loads(open('gpurun_out/nvls_${nv}_$1_$2_${N}gpu.json').read().splitlines()[-1]); print('nvls=$nv $1 $2 N=$N', 'ms/step %.4f' % d['ms_per_step'], 'value %.0f' % d['value'], {n: round(v['ms_per_step']*1e3,1) for n,v in d['kernels'].items() if 'adam' in n})" 2>&1 | tail -1
  done
done
echo "== done"
