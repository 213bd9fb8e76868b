#!/bin/bash
# Round 2, visit 3: first run of the fp32-parity tensor-core engine (CODAE_F32X3) on hardware.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
pick() { python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().splitlines()[-1]); k=d['kernels']; print(sys.argv[1], d['engine'][:24], 'ms/step %.4f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], {n: round(v['ms_per_step']*1e3,1) for n,v in k.items()})" $1 2>&1 | tail -1; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
echo "== x3 kernel tests"
timeout -s KILL 400 python -m pytest tests/test_gpu_f32x3.py -q -m gpu -p no:cacheprovider > gpurun_out/v3_x3_tests.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/v3_x3_tests.log
echo "== training tests"
timeout -s KILL 600 python -m pytest tests/test_gpu_training.py -q -m gpu -p no:cacheprovider > gpurun_out/v3_training.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/v3_training.log
E="python bench.py --workload embedding --steps 1000 --warmup 50 --no-cpu --no-scoring"
echo "== embedding fp32 (x3)"; timeout -s KILL 120 $E --dtype fp32 > gpurun_out/v3_emb_x3.json 2> gpurun_out/v3_emb_x3.err; echo "rc=$?"; pick gpurun_out/v3_emb_x3.json; tail -3 gpurun_out/v3_emb_x3.err
echo "== embedding fp32_simt"; timeout -s KILL 120 $E --dtype fp32_simt --steps 200 > gpurun_out/v3_emb_simt.json 2> gpurun_out/v3_emb_simt.err; echo "rc=$?"; pick gpurun_out/v3_emb_simt.json
echo "== embedding bf16"; timeout -s KILL 120 $E --dtype bf16 > gpurun_out/v3_emb_bf16.json 2> gpurun_out/v3_emb_bf16.err; echo "rc=$?"; pick gpurun_out/v3_emb_bf16.json
echo "== rest of the gpu suite"
timeout -s KILL 900 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_f32x3.py --deselect tests/test_gpu_training.py > gpurun_out/v3_suite.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/v3_suite.log
echo "== done"
