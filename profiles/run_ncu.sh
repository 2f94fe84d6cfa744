#!/bin/bash
# ncu evidence for the hot kernels (run under gpurun, 1 GPU). Outputs go to gpurun_out/; summaries are written
# into profiles/ by `python profiles/summarise.py <tag>` (plus opmix.py / stallmix.py on the source page).
#   bash profiles/run_ncu.sh <tag> [launches]     "launches" also records the per-launch duration list (slow)
#   KERNEL=tensor-full bash profiles/run_ncu.sh <tag>   profiles the tcgen05 family (default: bench's auto selection)
set -x
TAG=${1:-r1}
KERNEL=${KERNEL:-auto}
if [ "$2" == "launches" ]; then
  # every launch of a short bench with its device time (cold-cache, serialised: compare SHARES)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --kernel $KERNEL --steps 2 --warmup 3 --skip-cpu --skip-rebuild --no-graph > gpurun_out/launches_$TAG.log 2>&1
fi
# full capture of the hot kernels (one forward + one backward launch, tiled or tensor family) and the grad reduction
ncu --set full --clock-control none --import-source on -k 'regex:(fast|tc)_.*_kernel' -s 6 -c 3 -f -o gpurun_out/prof_$TAG \
    python bench.py --kernel $KERNEL --steps 2 --warmup 3 --skip-cpu --skip-rebuild --no-graph > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out | tail -5
