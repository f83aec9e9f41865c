#!/usr/bin/env bash
# Round-end, second call: whole GPU suite again (after the f1 test fix), then ncu --set full captures (source view) of the shipped
# attention kernel and of the opt-in tcgen05 one, for next round's attention work.  Usage (under gpurun): bash tools/gpu_r3b.sh [tag]
tag="${1:-r3b}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${tag}.log 2>&1
t0=$SECONDS
timeout 300 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit=$? after $((SECONDS - t0)) s"; tail -6 gpurun_out/pytest_gpu_${tag}.log
t0=$SECONDS
timeout 150 ncu --set full --clock-control none --import-source on -k regex:attention_head -s 3 -c 1 \
    -o gpurun_out/prof_attention_head_${tag} -f python bench.py --batch 64 --num-inference-steps 1 --steps 1 --warmup 1 \
    --no-cpu-baseline > gpurun_out/ncu_attention_head_${tag}.log 2>&1
echo "ncu v3 exit=$? after $((SECONDS - t0)) s"; tail -2 gpurun_out/ncu_attention_head_${tag}.log
t0=$SECONDS
PHENDIFF_B200_ATTN_KERNEL=tc timeout 150 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 3 -c 1 \
    -o gpurun_out/prof_attention_tc_${tag} -f python bench.py --batch 64 --num-inference-steps 1 --steps 1 --warmup 1 \
    --no-cpu-baseline > gpurun_out/ncu_attention_tc_${tag}.log 2>&1
echo "ncu tc exit=$? after $((SECONDS - t0)) s"; tail -2 gpurun_out/ncu_attention_tc_${tag}.log
ls -la gpurun_out | tail -6
