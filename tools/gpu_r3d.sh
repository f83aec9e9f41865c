#!/usr/bin/env bash
# Whole GPU suite, then short same-box benches of the attention variants inside the real forward (micro-batch 64).
# Usage (under gpurun): VARIANTS="tc:2 tc:3 tc:4 v3:6" bash tools/gpu_r3d.sh [tag]
tag="${1:-r3d}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${tag}.log 2>&1
t0=$SECONDS
timeout 200 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit=$? after $((SECONDS - t0)) s"; tail -8 gpurun_out/pytest_gpu_${tag}.log
for v in ${VARIANTS:-v3:6 tc:4 tc2:4 tc2:3 v3:6}; do
  k="${v%%:*}"; pp="${v##*:}"
  PHENDIFF_B200_ATTN_KERNEL=$k PHENDIFF_B200_ATTN_POLYPAIRS=$pp timeout 120 python bench.py --batch 64 --num-inference-steps 6 --steps 2 --warmup 3 \
      --no-cpu-baseline --e2e-steps 1 --dump-ops gpurun_out/ops_${tag}_${k}_$pp.md > gpurun_out/bench_${tag}_${k}_$pp.json 2> gpurun_out/bench_${tag}_${k}_$pp.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_${k}_$pp.json").read().strip().splitlines()[-1]); r=d["roofline"]
    att=[l for l in open("gpurun_out/ops_${tag}_${k}_$pp.md") if "attention S=" in l]
    print("variant $v", round(d["value"],2), "img/s (6+6 steps, batch 64)", "attn share", round(r["share_by_class"]["attention"],3), "attn ms", att[0].split("|")[4].strip() if att else None, d["clocks"]["sm_mhz"])
except Exception as e:
    print("variant $v FAILED", e)
PY
done
