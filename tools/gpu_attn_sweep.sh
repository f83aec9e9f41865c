#!/usr/bin/env bash
# Attention kernel parity + kernel / exp-mix sweep inside the real forward.  Usage (under gpurun): bash tools/gpu_attn_sweep.sh <tag>
# VARIANTS entries: <kernel>:<polypairs>, kernel in {tc, v3, v2}
tag="${1:-att}"
mkdir -p gpurun_out
log="gpurun_out/attn_${tag}.log"
: > "$log"
python -c "import __graft_entry__ as g; g.build()" >> "$log" 2>&1
timeout 600 python -m pytest -q --tb=short -p no:cacheprovider tests/test_gpu_kernels.py -m gpu -k "attention" >> "$log" 2>&1
echo "attention tests exit=$?" | tee -a "$log"
if [ -z "$SKIP_UNET" ]; then
timeout 900 python -m pytest -q --tb=short -p no:cacheprovider -s tests/test_gpu_unet.py -m gpu >> "$log" 2>&1
echo "unet tests exit=$?" | tee -a "$log"
fi
grep -E "passed|failed|FAILED|\[ddib|\[fp16 fwd\] small" "$log" | tail -20
# the first bench of a fresh box runs slow: throw one away
python bench.py --batch 64 --num-inference-steps 4 --steps 1 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
for v in ${VARIANTS:-tc:4 v3:6}; do
  k="${v%%:*}"; pp="${v##*:}"
  export PHENDIFF_B200_ATTN_KERNEL=$k PHENDIFF_B200_ATTN_POLYPAIRS=$pp
  timeout 600 python bench.py --batch 64 --num-inference-steps 6 --steps 2 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_${tag}_${k}_$pp.md \
      > gpurun_out/bench_${tag}_${k}_$pp.json 2>> "$log"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_${k}_$pp.json")); r=d["roofline"]
att=[l for l in open("gpurun_out/ops_${tag}_${k}_$pp.md") if "attention S=" in l]
print("variant $v", round(d["value"],2), "img/s(6 steps)", "attn share", round(r["share_by_class"]["attention"],3), "attn ms", att[0].split("|")[4].strip() if att else None)
PY
done
