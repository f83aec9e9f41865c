#!/usr/bin/env bash
# Full bench line + ncu launch list + ncu full capture of the dominant kernel.  Usage: bash tools/gpu_bench_profile.sh <tag>
tag="${1:-r1}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${tag}.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_${tag}.csv &
SMI=$!
python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
kill $SMI
tail -c 2500 gpurun_out/bench_${tag}.json
# launch list: one forward-sized window after the warm-up of a short run of the same workload (cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-3000} -c 330 --csv \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --batch 64 --num-inference-steps 10 --steps 1 --warmup 1 \
    --no-cpu-baseline > gpurun_out/ncu_launches_${tag}.log 2>&1
# DRAM traffic per launch over two whole forwards at the bench's micro-batch (roofline.traffic)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s ${NCU_SKIP_DRAM:-900} -c 202 --csv \
    --log-file gpurun_out/dram_${tag}.csv python bench.py --batch 64 --num-inference-steps 3 --steps 1 --warmup 1 \
    --no-cpu-baseline > gpurun_out/ncu_dram_${tag}.log 2>&1
# full capture of the dominant kernel (3 launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 60 -c 4 \
    -o gpurun_out/prof_conv_halo_${tag} -f python bench.py --batch 64 --num-inference-steps 2 --steps 1 --warmup 1 \
    --no-cpu-baseline > gpurun_out/ncu_full_${tag}.log 2>&1
ls -la gpurun_out | tail -12
