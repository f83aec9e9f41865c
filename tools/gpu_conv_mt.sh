#!/usr/bin/env bash
# Dual-accumulator (16x16-pixel) halo tiles: parity + A/B against 16x8 tiles.  Usage (under gpurun): bash tools/gpu_conv_mt.sh <tag>
tag="${1:-mt}"
mkdir -p gpurun_out
log="gpurun_out/convmt_${tag}.log"
: > "$log"
python -c "import __graft_entry__ as g; g.build()" >> "$log" 2>&1
timeout 600 python -m pytest -q --tb=short -p no:cacheprovider tests/test_gpu_kernels.py -m gpu -k "halo or chunk_statistics or conv_out_ddim or tcgen05" >> "$log" 2>&1
echo "conv kernel tests exit=$?" | tee -a "$log"
timeout 900 python -m pytest -q --tb=short -p no:cacheprovider -s tests/test_gpu_unet.py -m gpu >> "$log" 2>&1
echo "unet tests exit=$?" | tee -a "$log"
grep -E "passed|failed|FAILED|\[ddib" "$log" | tail -12
for v in ${VARIANTS:-1 2}; do
  export PHENDIFF_B200_HALO_MT=$v
  timeout 600 python bench.py --batch 64 --num-inference-steps 6 --steps 2 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_${tag}_mt$v.md \
      > gpurun_out/bench_${tag}_mt$v.json 2>> "$log"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_mt$v.json")); r=d["roofline"]
print("MT=$v", round(d["value"],2), "img/s(6 steps)  conv TF", round(r["achieved"],1), {k:round(x,3) for k,x in r["share_by_class"].items() if x>0.005})
PY
done
