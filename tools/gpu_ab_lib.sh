#!/usr/bin/env bash
# Same-box A/B of two builds of the library (short benches, alternating).  Usage: bash tools/gpu_ab_lib.sh <tag> <libA> <libB> [rounds]
tag="$1"; A="$2"; B="$3"; R="${4:-2}"
mkdir -p gpurun_out
for r in $(seq 1 $R); do
  for v in A B; do
    lib=$([ $v = A ] && echo "$A" || echo "$B")
    PHENDIFF_B200_LIB="$PWD/$lib" timeout 600 python bench.py --num-inference-steps 10 --steps 2 --warmup 1 --no-cpu-baseline \
        --dump-ops gpurun_out/ops_${tag}_${v}${r}.md > gpurun_out/bench_${tag}_${v}${r}.json 2> gpurun_out/bench_${tag}_${v}${r}.err
    python - "$v$r" "$lib" gpurun_out/bench_${tag}_${v}${r}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1]); r = d["roofline"]
    print(sys.argv[1], sys.argv[2], round(d["value"], 2), "img/s; e2e", round(d["e2e"]["value"], 2), "conv TF", round(r["achieved"]), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
