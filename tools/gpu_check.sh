#!/bin/bash
# One GPU-box visit: probe, parity tests (separate processes so a CUDA fault in one cannot poison the rest),
# smoke, bench lines, and the ncu launch list of a short bench run.  Everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== sanity"; timeout -s KILL 120 python -c "import sys; sys.path.insert(0, 'mui-deepautoencoder_b200'); import torch; from codae import _C; print(_C.lib().codae_ctx_sm_count(_C.ctx(torch.device('cuda', 0))), 'SMs')" 2>&1 | tail -2
echo "== probe"; timeout -s KILL 150 python tests/gpu_probe_gemm.py > gpurun_out/probe.log 2>&1; prc=$?; echo "probe rc=$prc"; tail -25 gpurun_out/probe.log
if [ $prc -eq 137 ]; then echo "probe hung: stopping this visit"; exit 3; fi
echo "== pytest kernels (no tcgen05)"; timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tcgen05" -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_kernels.log
echo "== pytest tcgen05"; timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tcgen05" -p no:cacheprovider > gpurun_out/pytest_tcgen05.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_tcgen05.log
echo "== pytest training"; timeout -s KILL 600 python -m pytest tests/test_gpu_training.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_training.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_training.log
echo "== pytest inference"; timeout -s KILL 600 python -m pytest tests/test_gpu_inference.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_inference.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_inference.log
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== bench fp32"; timeout -s KILL 900 python bench.py > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "rc=$?"; tail -3 gpurun_out/bench_fp32.err; cat gpurun_out/bench_fp32.json
echo "== bench bf16"; timeout -s KILL 900 python bench.py --dtype bf16 --no-cpu > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "rc=$?"; tail -3 gpurun_out/bench_bf16.err; cat gpurun_out/bench_bf16.json
echo "== reference arm"; timeout -s KILL 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
echo "== ncu launch list"
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-scoring --no-cpu --no-graph > gpurun_out/plain.log 2>&1 && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-scoring --no-cpu --no-graph > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
echo "== done"
