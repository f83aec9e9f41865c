#!/usr/bin/env bash
# Quick GPU iteration: tensor-core kernel tests + UNet parity + a reduced-workload bench sweep.  Usage: bash tools/gpu_quick.sh <tag>
tag="${1:-q}"
mkdir -p gpurun_out
log="gpurun_out/quick_${tag}.log"
: > "$log"
python -c "import __graft_entry__ as g; g.build()" >> "$log" 2>&1
run() {
  echo "=== $* ===" | tee -a "$log"
  timeout 900 python -m pytest -q --tb=short -p no:cacheprovider -s "$@" >> "$log" 2>&1
  echo "exit=$?" | tee -a "$log"
}
run tests/test_gpu_kernels.py -m gpu -k "${KSEL:-tcgen05 or halo or chunk_statistics or conv_out_ddim or groupnorm or attention}"
run tests/test_gpu_unet.py -m gpu
grep -E "^(=== |exit=|FAILED|ERROR)|passed|failed|\[ddib|\[bf16 fwd\] small|\[fp16 fwd\] small" "$log" | tail -40
MBS="${MBS:-32}" bash tools/gpu_sweep_mb.sh ${tag}
