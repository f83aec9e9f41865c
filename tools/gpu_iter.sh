#!/usr/bin/env bash
# One GPU iteration: selected kernel tests, UNet parity, full bench line.  Usage: bash tools/gpu_iter.sh <tag> ["<pytest -k expr>"]
tag="${1:-it}"
ksel="${2:-fused_groupnorm or halo or chunk_statistics or conv_out_ddim}"
mkdir -p gpurun_out
log="gpurun_out/iter_${tag}.log"
: > "$log"
python -c "import __graft_entry__ as g; g.build()" >> "$log" 2>&1
run() {
  echo "=== $* ===" | tee -a "$log"
  timeout 600 python -m pytest -q --tb=short -p no:cacheprovider -s "$@" >> "$log" 2>&1
  echo "exit=$?" | tee -a "$log"
}
run tests/test_gpu_kernels.py -m gpu -k "$ksel"
run tests/test_gpu_unet.py -m gpu
grep -E "^(=== |exit=|FAILED|ERROR)|passed|failed|Error|error|\[ddib|\[bf16 fwd\] small|\[fp16 fwd\] small" "$log" | tail -40
if [ -z "${NO_BENCH:-}" ]; then
  timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
  tail -c 2600 gpurun_out/bench_${tag}.json
  tail -5 gpurun_out/bench_${tag}.err
fi
