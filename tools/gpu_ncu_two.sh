#!/usr/bin/env bash
# ncu --set full captures of two named kernels at micro-batch 64 (the bench's), plus an N=2 torchrun bench when 2 GPUs are visible
# Usage: bash tools/gpu_ncu_two.sh <tag> '<regex1>' '<regex2>'
tag="${1:-n2}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${tag}.log 2>&1
i=0
for rx in "$2" "$3"; do
  i=$((i+1))
  [ -z "$rx" ] && continue
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$rx" -s ${NCU_SKIP_K:-6} -c 2 \
      -o gpurun_out/prof_k${i}_${tag} -f python bench.py --batch 64 --num-inference-steps 2 --steps 1 --warmup 1 \
      --no-cpu-baseline > gpurun_out/ncu_k${i}_${tag}.log 2>&1
  tail -3 gpurun_out/ncu_k${i}_${tag}.log
done
ls -la gpurun_out | tail -6
