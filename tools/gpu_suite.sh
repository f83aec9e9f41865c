#!/usr/bin/env bash
# Whole GPU suite in one process, log into gpurun_out/.  Usage (under gpurun): bash tools/gpu_suite.sh [tag]
tag="${1:-suite}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${tag}.log 2>&1
t0=$SECONDS
timeout 300 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit=$? after $((SECONDS - t0)) s"; tail -30 gpurun_out/pytest_gpu_${tag}.log
