#!/usr/bin/env bash
# Run the GPU test groups in separate processes (a trapped kernel poisons its CUDA context, not the next group).
# Usage (under gpurun): bash tools/gpu_check.sh [tag]
tag="${1:-r1}"
mkdir -p gpurun_out
log="gpurun_out/gpu_check_${tag}.log"
: > "$log"
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> "$log" 2>&1
python -c "import __graft_entry__ as g; g.build()" >> "$log" 2>&1
run() {
  echo "=== $* ===" | tee -a "$log"
  timeout 900 python -m pytest -q --tb=short -p no:cacheprovider -s "$@" >> "$log" 2>&1
  echo "exit=$?" | tee -a "$log"
}
run tests/test_gpu_kernels.py -m gpu -k "conv_simt"
run tests/test_gpu_kernels.py -m gpu -k "groupnorm or scheduler or add_noise"
run tests/test_gpu_kernels.py -m gpu -k "attention"
run tests/test_gpu_kernels.py -m gpu -k "tcgen05 or halo or chunk_statistics or conv_out_ddim"
run tests/test_gpu_unet.py -m gpu -k "fp32 or api or fused or pipeline or golden"
run tests/test_gpu_unet.py -m gpu -k "half"
grep -E "^(=== |exit=|FAILED|ERROR|[0-9]+ (passed|failed))|passed|failed|\[ddib|\[bf16|\[fp16" "$log" | tail -80
