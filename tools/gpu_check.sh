#!/bin/bash
# One GPU-box visit: probe, parity tests (separate processes so a CUDA fault in one cannot poison the rest),
# smoke, bench lines, and the ncu launch list / full capture of a short bench run.  Everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
STAGES="${1:-all}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== sanity"; timeout -s KILL 120 python -c "import sys; sys.path.insert(0, 'mui-deepautoencoder_b200'); import torch; from codae import _C; print(_C.lib().codae_ctx_sm_count(_C.ctx(torch.device('cuda', 0))), 'SMs')" 2>&1 | tail -2
echo "== probe"; timeout -s KILL 150 python tests/gpu_probe_gemm.py > gpurun_out/probe.log 2>&1; prc=$?; echo "probe rc=$prc"; grep -c "bad=0/" gpurun_out/probe.log; grep -v "bad=0/" gpurun_out/probe.log | tail -12
if [ $prc -eq 137 ]; then echo "probe hung: stopping this visit"; exit 3; fi
run_pytest() { echo "== pytest $1"; timeout -s KILL 600 python -m pytest $2 -m gpu -q -p no:cacheprovider > gpurun_out/pytest_$1.log 2>&1; echo "rc=$?"; tail -${3:-15} gpurun_out/pytest_$1.log; }
if [ "${SPLIT_PYTEST:-0}" = "1" ]; then
run_pytest kernels tests/test_gpu_kernels.py 40
run_pytest training tests/test_gpu_training.py 40
run_pytest inference tests/test_gpu_inference.py 15
run_pytest scripts tests/test_gpu_scripts.py 15
else
# exactly what the driver runs at round end: the whole GPU suite in one process
echo "== pytest (driver command)"; timeout -s KILL 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
fi
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench default"; timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
echo "== bench modanet"; timeout -s KILL 300 python bench.py --workload modanet --no-cpu --no-scoring --no-fp32 > gpurun_out/bench_modanet.json 2> gpurun_out/bench_modanet.err; echo "rc=$?"; tail -3 gpurun_out/bench_modanet.err; cat gpurun_out/bench_modanet.json
echo "== bench polyvore"; timeout -s KILL 300 python bench.py --workload polyvore --steps 10 --warmup 3 --no-cpu --no-scoring > gpurun_out/bench_polyvore.json 2> gpurun_out/bench_polyvore.err; echo "rc=$?"; tail -3 gpurun_out/bench_polyvore.err; cat gpurun_out/bench_polyvore.json
echo "== reference arm"; timeout -s KILL 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
echo "== ncu launch list (default command, short)"
CMD="python bench.py --steps 3 --warmup 3 --no-scoring --no-cpu --no-fp32 --no-graph"
timeout -s KILL 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
echo "== ncu full capture of the hot kernels"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu --no-fp32 --no-graph --catalog 2000000"
timeout -s KILL 300 $CMD2 > gpurun_out/plain2.log 2>&1 && {
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"adam_kernel|adam_partials_kernel|sqnorm_kernel|corrupt_fwd|mse_loss" -s 9 -c 6 -o gpurun_out/prof_r1_elementwise $CMD2 > gpurun_out/ncu_full1.log 2>&1; echo "ncu elementwise rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"score_topk_kernel" -s 2 -c 2 -o gpurun_out/prof_r1_score $CMD2 > gpurun_out/ncu_full2.log 2>&1; echo "ncu score rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"tc05_gemm_kernel" -s 100 -c 8 -o gpurun_out/prof_r1_gemm_small $CMD2 > gpurun_out/ncu_full3.log 2>&1; echo "ncu gemm small rc=$?"
}
CMD3="python bench.py --workload polyvore --steps 1 --warmup 3 --no-cpu --no-scoring --no-graph"
timeout -s KILL 300 $CMD3 > gpurun_out/plain3.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"tc05_gemm" -s 96 -c 8 -o gpurun_out/prof_r1_gemm_large $CMD3 > gpurun_out/ncu_full4.log 2>&1
echo "ncu gemm large rc=$?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"corrupt_fwd|mse_loss|clip_adam_kernel|adam_partials_kernel" -s 6 -c 3 -o gpurun_out/prof_r1_elementwise_large $CMD3 > gpurun_out/ncu_full5.log 2>&1
echo "ncu elementwise large rc=$?"
echo "== done"
