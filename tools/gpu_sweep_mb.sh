#!/usr/bin/env bash
# Micro-batch sweep of the whole path on a reduced workload (same per-forward shapes; fewer steps).  Usage: bash tools/gpu_sweep_mb.sh <tag>
tag="${1:-sweep}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${tag}.log 2>&1
: > gpurun_out/sweep_${tag}.jsonl
for mb in ${MBS:-4 8 16 32 64}; do
  python bench.py --batch 64 --num-inference-steps 10 --steps 2 --warmup 3 --microbatch $mb --no-cpu-baseline --dump-ops gpurun_out/ops_${tag}_mb$mb.md \
      >> gpurun_out/sweep_${tag}.jsonl 2>> gpurun_out/sweep_${tag}.err
done
python - <<PY
import json
for l in open("gpurun_out/sweep_${tag}.jsonl"):
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["microbatch"], round(d["value"],2), "img/s  conv TF", round(r["achieved"],1), {k:round(v,3) for k,v in r["share_by_class"].items()})
PY
