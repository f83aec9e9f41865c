#!/usr/bin/env bash
# Probe the UMMA descriptor semantics the halo kernel relies on (shifted start inside a SWIZZLE_128B tile): run its parity
# tests under the descriptor variants.  Usage (under gpurun): bash tools/gpu_halo_probe.sh <tag>
tag="${1:-probe}"
mkdir -p gpurun_out
log="gpurun_out/halo_probe_${tag}.log"
: > "$log"
python -c "import __graft_entry__ as g; g.build()" >> "$log" 2>&1
run() {
  echo "=== PITCH=$1 BASEOFF=$2 ===" | tee -a "$log"
  PHENDIFF_B200_HALO_PITCH=$1 PHENDIFF_B200_HALO_BASEOFF=$2 timeout 600 python -m pytest -q --tb=line -p no:cacheprovider \
      tests/test_gpu_kernels.py -m gpu -k "halo or chunk_statistics or conv_out_ddim" >> "$log" 2>&1
  echo "exit=$?" | tee -a "$log"
}
run tight 1
run tight 0
run pow2 1
run pow2 0
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/mufu tools/microbench/mufu.cu && ./gpurun_out/mufu | tee gpurun_out/mufu_${tag}.txt
rm -f gpurun_out/mufu
grep -E "^(=== |exit=)|passed|failed" "$log"
grep -E "^(FAILED|ERROR|/root|tests/)" "$log" | head -60
