#!/usr/bin/env bash
# A/B of the GroupNorm-in-conv fusion: kernel tests in both SiLU modes, then short benches (10+10 steps) with per-op tables.
tag="${1:-ab}"
mkdir -p gpurun_out
log="gpurun_out/gnab_${tag}.log"
: > "$log"
python -c "import __graft_entry__ as g; g.build()" >> "$log" 2>&1
for silu in exp tanh; do
  echo "=== kernel tests, GN_SILU=$silu ===" | tee -a "$log"
  PHENDIFF_B200_GN_SILU=$silu timeout 600 python -m pytest -q --tb=short -p no:cacheprovider -s tests/test_gpu_kernels.py -m gpu -k "fused_groupnorm" >> "$log" 2>&1
  echo "exit=$?" | tee -a "$log"
done
grep -E "passed|failed|FAILED|max abs err" "$log" | tail -30
bench() {  # name, env...
  local name="$1"; shift
  env "$@" timeout 600 python bench.py --num-inference-steps 10 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 \
      --dump-ops gpurun_out/ops_${tag}_${name}.md > gpurun_out/bench_${tag}_${name}.json 2> gpurun_out/bench_${tag}_${name}.err
  python - "$name" gpurun_out/bench_${tag}_${name}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{sys.argv[1]:>14}: {d['value']:.2f} img/s  conv {r['achieved']:.0f} TF/s  shares {{k: round(v, 3) for k, v in r['share_by_class'].items() if v > 0.005}}  clk {d['clocks']['sm_mhz']}".replace("{{", "{").replace("}}", "}"))
    print("   ", {k: round(v, 3) for k, v in r["share_by_class"].items() if v > 0.005})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
bench nofuse PHENDIFF_B200_GNFUSE=0
bench exp PHENDIFF_B200_GN_SILU=exp
bench tanh PHENDIFF_B200_GN_SILU=tanh
bench tanh_sa2 PHENDIFF_B200_GN_SILU=tanh PHENDIFF_B200_HALO_GN_SA=2
bench exp_sa2 PHENDIFF_B200_GN_SILU=exp PHENDIFF_B200_HALO_GN_SA=2
