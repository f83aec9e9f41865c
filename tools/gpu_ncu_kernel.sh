#!/usr/bin/env bash
# ncu --set full capture of one kernel class inside a short run of the bench workload.
# Usage (under gpurun): bash tools/gpu_ncu_kernel.sh <tag> <kernel-regex> [skip] [count]
tag="${1:-r1}"; pat="${2:-attention}"; skip="${3:-4}"; cnt="${4:-3}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${tag}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${pat} -s ${skip} -c ${cnt} \
    -o gpurun_out/prof_${pat}_${tag} -f python bench.py --batch 32 --num-inference-steps 2 --steps 1 --warmup 1 \
    --no-cpu-baseline > gpurun_out/ncu_full_${pat}_${tag}.log 2>&1
tail -3 gpurun_out/ncu_full_${pat}_${tag}.log
ls -la gpurun_out | tail -5
