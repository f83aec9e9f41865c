#!/usr/bin/env bash
tag=${1:-r2a}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_$tag.log 2>&1
timeout 600 python -m pytest -q --tb=short -p no:cacheprovider tests/test_gpu_kernels.py -m gpu -k "fused_groupnorm or halo" 2>&1 | tail -3
timeout 600 python bench.py --num-inference-steps 10 --steps 2 --warmup 1 --no-cpu-baseline --dump-ops gpurun_out/ops_$tag.md > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - $tag <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1]); r = d["roofline"]
print(round(d["value"], 2), "img/s (10+10) conv TF", round(r["achieved"]), {k: round(v, 3) for k, v in r["share_by_class"].items() if v > 0.005}, d["clocks"]["sm_mhz"])
PY
NCU_SKIP_K=6 bash tools/gpu_ncu_two.sh $tag 'conv_halo_kernelILi128E6__halfLi0ELi4ELi2ELb1E' ''
