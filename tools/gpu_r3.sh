#!/usr/bin/env bash
# Round-end confirmation on one B200: whole GPU suite (incl. the published-KAT scheduler tests and the f1 drop-in), the secondary
# CFG bench line, then the default bench line.  Every stage has its own timeout and writes to gpurun_out/ as it goes.
# Usage (under gpurun): bash tools/gpu_r3.sh [tag]
tag="${1:-r3}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build_${tag}.log 2>&1
t0=$SECONDS
timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu_${tag}.log 2>&1
echo "pytest exit=$? after $((SECONDS - t0)) s"; tail -14 gpurun_out/pytest_gpu_${tag}.log
t0=$SECONDS
timeout 200 python bench.py --workload cfg --steps 1 --warmup 3 > gpurun_out/bench_cfg_${tag}.json 2> gpurun_out/bench_cfg_${tag}.err
echo "cfg bench exit=$? after $((SECONDS - t0)) s"; cut -c1-400 gpurun_out/bench_cfg_${tag}.json; tail -3 gpurun_out/bench_cfg_${tag}.err
t0=$SECONDS
timeout 400 python bench.py --dump-ops gpurun_out/ops_${tag}.md > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench exit=$? after $((SECONDS - t0)) s"; cut -c1-300 gpurun_out/bench_${tag}.json; tail -3 gpurun_out/bench_${tag}.err
