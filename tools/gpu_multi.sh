#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL parity check, then bench at N (and optionally smaller N) GPUs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
LIST=${2:-"1 2 4 8"}
nvidia-smi -L | head -8
echo "== dist check N=$N"
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "rc=$?"; grep -E "DP |sharded|DIST CHECK|Error|error" gpurun_out/dist_check_$N.log | tail -12
for n in $LIST; do
  if [ $n -le $N ]; then
    echo "== bench N=$n"
    if [ $n -eq 1 ]; then
      timeout -s KILL 240 python bench.py --gpus 1 --no-cpu --no-fp32 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    else
      timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu --no-fp32 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    fi
    echo "rc=$?"; tail -2 gpurun_out/scale_$n.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_$n.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=%d value %.0f samples/s  ms/step %.4f  e2e %.0f  scoring %.3g scores/s (%.3f ms/sweep)"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["scoring"]["value"],d["scoring"]["ms_per_sweep"]))
except Exception as e: print("parse error", e)
PY
  fi
done
echo "== done"
