#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CODAE_EXPERIMENTAL=1 timeout -s KILL 16 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_experimental.py -k "persistent or layerwise" -q -m gpu -p no:cacheprovider > gpurun_out/tiny.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/tiny.log
