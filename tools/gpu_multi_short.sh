#!/bin/bash
# Short multi-GPU visit (gpurun --gpus N): NCCL parity check, swap-mode inference under torchrun, bench at N GPUs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
echo "== dist check N=$N"
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "rc=$?"; grep -E "DP |sharded|DIST CHECK|Error|error" gpurun_out/dist_check_$N.log | tail -12
cd mui-deepautoencoder_b200
for mode in slot swap; do
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 script/4_complementarity_inference.py --config config/embedding.yaml --synthetic 20000 --slot 2 --k 5 --queries 2 --mode $mode > ../gpurun_out/script_infer_${mode}_$N.log 2>&1; echo "infer $mode N=$N rc=$?"; grep -E "indices|Error" ../gpurun_out/script_infer_${mode}_$N.log | cut -c1-220 | tail -2
timeout -s KILL 120 python script/4_complementarity_inference.py --config config/embedding.yaml --synthetic 20000 --slot 2 --k 5 --queries 2 --mode $mode 2>/dev/null | grep indices | cut -c1-220
done
cd ..
echo "== bench N=$N"
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu --no-fp32 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
echo "rc=$?"; tail -2 gpurun_out/scale_$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_$N.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=%d value %.0f samples/s  ms/step %.4f  e2e %.0f  scoring %.3g scores/s (%.3f ms/sweep) swaps %.3g/s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["scoring"]["value"],d["scoring"]["ms_per_sweep"],d["scoring"]["swap_reconstruction"]["value"]))
except Exception as e: print("parse error", e)
PY
echo "== done"
