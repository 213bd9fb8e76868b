#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL parity check, then bench at N (and optionally smaller N) GPUs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
LIST=${2:-"1 2 4 8"}
nvidia-smi -L | head -8
echo "== dist check N=$N"
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "rc=$?"; grep -E "DP |sharded|DIST CHECK|Error|error" gpurun_out/dist_check_$N.log | tail -12
echo "== scripts under torchrun N=$N"
cd mui-deepautoencoder_b200
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 script/train_dae_on_embedding.py --output_path /tmp/out --config config/embedding.yaml --synthetic 1000 --epochs 2 --graph > ../gpurun_out/script_train_$N.log 2>&1; echo "train rc=$?"; grep -E "TRAINING FULL ERROR|VALIDATION FULL ERROR|ENDED|Error" ../gpurun_out/script_train_$N.log | tail -6
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 script/4_complementarity_inference.py --config config/embedding.yaml --synthetic 20000 --slot 2 --k 5 --queries 2 > ../gpurun_out/script_infer_$N.log 2>&1; echo "infer rc=$?"; grep -E "indices|Error" ../gpurun_out/script_infer_$N.log | cut -c1-200 | tail -2
timeout -s KILL 120 python script/4_complementarity_inference.py --config config/embedding.yaml --synthetic 20000 --slot 2 --k 5 --queries 2 2>/dev/null | grep indices | cut -c1-200
cd ..
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
