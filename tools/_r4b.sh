mkdir -p gpurun_out
for sw in 0 1 2 3; do
  echo "=== tc3 parity, SWAP=$sw"
  PHENDIFF_B200_ATTN_TC3_SWAP=$sw timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --tb=line -k "mmatc3" -x 2>&1 | tail -5
done
bash tools/gpu.sh r4b "pytest:guided or groupnorm or headline_forward or san_ or cfg or classifier" micro
bash tools/gpu.sh r4b "attn:v3:6 tc3:4 tc3:3 tc3:2 tc3:5"
