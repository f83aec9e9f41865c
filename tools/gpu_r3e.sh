#!/usr/bin/env bash
# Published-vector GPU tests + an ncu capture of the reworked tcgen05 attention kernel.  Usage (under gpurun): bash tools/gpu_r3e.sh [tag]
tag="${1:-r3e}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${tag}.log 2>&1
timeout 120 python -m pytest tests/test_gpu_published_kats.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/pytest_kats_${tag}.log 2>&1
echo "pytest exit=$?"; tail -25 gpurun_out/pytest_kats_${tag}.log
PHENDIFF_B200_ATTN_KERNEL=tc timeout 100 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 3 -c 1 \
    -o gpurun_out/prof_attention_tc_${tag} -f python bench.py --batch 64 --num-inference-steps 1 --steps 1 --warmup 1 \
    --no-cpu-baseline > gpurun_out/ncu_attention_tc_${tag}.log 2>&1
echo "ncu tc exit=$?"; tail -2 gpurun_out/ncu_attention_tc_${tag}.log
