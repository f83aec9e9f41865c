#!/usr/bin/env bash
# Attention kernel parity + exp-mix sweep inside the real forward.  Usage (under gpurun): bash tools/gpu_attn_sweep.sh <tag>
tag="${1:-att}"
mkdir -p gpurun_out
log="gpurun_out/attn_${tag}.log"
: > "$log"
python -c "import __graft_entry__ as g; g.build()" >> "$log" 2>&1
timeout 600 python -m pytest -q --tb=short -p no:cacheprovider tests/test_gpu_kernels.py -m gpu -k "attention" >> "$log" 2>&1
echo "attention tests exit=$?" | tee -a "$log"
timeout 900 python -m pytest -q --tb=short -p no:cacheprovider -s tests/test_gpu_unet.py -m gpu >> "$log" 2>&1
echo "unet tests exit=$?" | tee -a "$log"
grep -E "passed|failed|\[ddib|\[bf16 fwd\]|\[fp16 fwd\]" "$log" | tail -20
for v in ${VARIANTS:-v2 0 4 6 7 8 10}; do
  if [ "$v" = "v2" ]; then export PHENDIFF_B200_ATTN_KERNEL=v2; unset PHENDIFF_B200_ATTN_POLYPAIRS; else unset PHENDIFF_B200_ATTN_KERNEL; export PHENDIFF_B200_ATTN_POLYPAIRS=$v; fi
  python bench.py --batch 64 --num-inference-steps 6 --steps 2 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_${tag}_$v.md \
      > gpurun_out/bench_${tag}_$v.json 2>> "$log"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_$v.json")); r=d["roofline"]
att=[l for l in open("gpurun_out/ops_${tag}_$v.md") if "attention S=" in l]
print("variant $v", round(d["value"],2), "img/s(6 steps)", "attn share", round(r["share_by_class"]["attention"],3), "attn ms", att[0].split("|")[4].strip() if att else None)
PY
done
