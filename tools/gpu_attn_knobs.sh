#!/usr/bin/env bash
# short benches over the attention kernel's polynomial-exp share (PHENDIFF_B200_ATTN_POLYPAIRS: score pairs of 16 on the FMA/ALU pipes)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_knobs.log 2>&1
for pp in ${PPS:-6 4 7 8 10 6}; do
  PHENDIFF_B200_ATTN_POLYPAIRS=$pp timeout 600 python bench.py --num-inference-steps 10 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_pp$pp.json 2> gpurun_out/bench_pp$pp.err
  python - $pp <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench_pp{sys.argv[1]}.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("polypairs", sys.argv[1], round(d["value"], 2), "img/s; e2e", round(d["e2e"]["value"], 2), "attention share", round(r["share_by_class"]["attention"], 4), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
