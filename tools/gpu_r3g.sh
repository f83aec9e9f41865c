#!/usr/bin/env bash
# Last call of round 1: the default attention tests (dispatch untouched for the shipped kernels), then the first ever run of the
# experimental tc2 kernel in its own process (a trap there must not poison the other tests).  Usage (under gpurun): bash tools/gpu_r3g.sh
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --tb=line -k "test_attention" > gpurun_out/pytest_attn_r3g.log 2>&1
echo "default attention tests exit=$?"; tail -4 gpurun_out/pytest_attn_r3g.log
PHENDIFF_B200_EXPERIMENTAL=1 timeout 30 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --tb=short -k "experimental_tc2" > gpurun_out/pytest_tc2_r3g.log 2>&1
echo "tc2 tests exit=$?"; tail -15 gpurun_out/pytest_tc2_r3g.log
