#!/bin/bash
# compute-sanitizer over the hot path (run under gpurun, 1 GPU): memcheck over smoke() (takes, window sums, tensor forward /
# reverse kernels with their mbarrier / TMA / tensor-memory paths, reduce kernels, Adam, CUDA-graph capture), racecheck
# (shared-memory hazards) over the torch-free C-ABI check of the tiled and tensor kernels.  Summaries: profiles/<tag>_sanitizer.md
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck.log 2>&1
timeout 700 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tests/tools/tc_capi_check.py cfg5 > gpurun_out/${TAG}_racecheck.log 2>&1
tail -n 3 gpurun_out/${TAG}_memcheck.log; tail -n 3 gpurun_out/${TAG}_racecheck.log
