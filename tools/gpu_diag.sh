#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 120 python tests/gpu_diag_tma_store.py > gpurun_out/diag_tma.log 2>&1; echo "rc=$?"; cat gpurun_out/diag_tma.log | tail -40
