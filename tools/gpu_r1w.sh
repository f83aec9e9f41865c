#!/usr/bin/env bash
# parity of the UNet / DDIB path in tanh-SiLU mode, then a micro-batch sweep of the fused path (short benches)
mkdir -p gpurun_out
log=gpurun_out/r1w.log; : > $log
python -c "import __graft_entry__ as g; g.build()" >> $log 2>&1
echo "=== unet parity, GN_SILU=tanh ===" | tee -a $log
PHENDIFF_B200_GN_SILU=tanh timeout 600 python -m pytest -q --tb=short -p no:cacheprovider -s tests/test_gpu_unet.py -m gpu -k "half or golden" >> $log 2>&1
echo "exit=$?" | tee -a $log
grep -E "passed|failed|FAILED|\[ddib|\[bf16 fwd\]|\[fp16 fwd\]" $log | tail -30
for mb in 32 64 128; do
  PHENDIFF_B200_GN_SILU=tanh timeout 600 python bench.py --num-inference-steps 10 --steps 2 --warmup 1 --no-cpu-baseline --microbatch $mb \
     --dump-ops gpurun_out/ops_r1w_mb$mb.md > gpurun_out/bench_r1w_mb$mb.json 2> gpurun_out/bench_r1w_mb$mb.err
  python - $mb <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench_r1w_mb{sys.argv[1]}.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("mb", sys.argv[1], round(d["value"], 2), "img/s conv TF", round(r["achieved"]), {k: round(v, 3) for k, v in r["share_by_class"].items() if v > 0.005}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("mb", sys.argv[1], "FAILED", e)
PY
done
