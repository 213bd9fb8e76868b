#!/bin/bash
# Round 2, visit 2 (1 GPU): whole GPU suite on the cleaned-up tree, smoke on the tensor-core engine, the new default bench line
# (polyvore-shaped primary + secondary blocks + scoring + live-reference CPU arm), abalone tiny-MLP A/B.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest (driver command)"; timeout -s KILL 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/v2_pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/v2_pytest_gpu.log
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v2_smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/v2_smoke.log
echo "== abalone A/B"; timeout -s KILL 200 python tools/abalone_ab.py 2>&1 | tail -6
echo "== bench default (driver command)"; ( time timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/v2_bench_default.json 2> gpurun_out/v2_bench_default.err ) 2>&1 | tail -3; echo "rc=$?"; tail -3 gpurun_out/v2_bench_default.err; cut -c1-1500 gpurun_out/v2_bench_default.json
echo "== reference arm (driver command)"; ( time timeout -s KILL 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/v2_bench_ref.json 2> gpurun_out/v2_bench_ref.err ) 2>&1 | tail -3; cut -c1-900 gpurun_out/v2_bench_ref.json
echo "== done"
