#!/usr/bin/env bash
# round-1 reference run of the fused path: full GPU test suite, bench line, ncu launch list + full capture of the GN-variant conv
tag=${1:-r2d}
mkdir -p gpurun_out
bash tools/gpu_check.sh $tag
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -2 gpurun_out/smoke_$tag.log
BENCH_ARGS="--dump-ops gpurun_out/ops_$tag.md" NCU_SKIP=4400 bash tools/gpu_bench_profile.sh $tag
