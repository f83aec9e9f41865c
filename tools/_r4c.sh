mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider --tb=line -k "mmatc3" -x 2>&1 | tail -5
bash tools/gpu.sh r4c "attn:v3:6 tc3:4 tc3:3 tc3:2 tc3:0"
