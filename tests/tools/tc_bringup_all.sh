#!/bin/bash
# One-call hardware bring-up of the tensor (tcgen05) family; run under gpurun on one B200:
#     gpurun --timeout 1500 -- 'bash tests/tools/tc_bringup_all.sh'
# Every stage has its own timeout and writes into gpurun_out/; a failing stage does not stop the later ones.
# Stages: 1 torch-free C-ABI check of every instance   2 oracle-level pytest of the opt-in variants
#         3 forward timings + breakdown (v1, pipelined v2)   4 reverse timings + breakdown
#         5 bench lines: auto, tensor-full, tensor-full with the pipelined forward   6 ncu capture of the tc_ kernels
set -x
mkdir -p gpurun_out
OUT=gpurun_out
timeout 120 python tests/tools/tc_selftest.py --mn                   > $OUT/tcall_0_selftest_mn.log 2>&1
timeout 120 python tests/tools/tc_capi_check.py all                 > $OUT/tcall_1_capi.log 2>&1
FBP_TC_TESTS=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu > $OUT/tcall_2_pytest.log 2>&1
FBP_ACT_TESTS=1 timeout 300 python -m pytest tests/test_gpu_networks.py -q -m gpu > $OUT/tcall_2b_networks.log 2>&1
timeout 300 python tests/tools/tc_bringup.py                         > $OUT/tcall_3_fwd.log 2>&1
timeout 300 python tests/tools/tc_bringup_bwd.py                     > $OUT/tcall_4_bwd.log 2>&1
timeout 300 python bench.py --skip-cpu --steps 100 --warmup 5                       > $OUT/tcall_5_bench_auto.json 2> $OUT/tcall_5_bench_auto.err
timeout 300 python bench.py --skip-cpu --steps 100 --warmup 5 --kernel tensor-full  > $OUT/tcall_5_bench_full.json 2> $OUT/tcall_5_bench_full.err
FBP_TC_FWD=2 timeout 300 python bench.py --skip-cpu --steps 100 --warmup 5 --kernel tensor-full > $OUT/tcall_5_bench_full_v2.json 2> $OUT/tcall_5_bench_full_v2.err
# the whole GPU suite with the tensor reverse kernel selected by auto mode (what flipping the default would run)
FBP_TC_AUTO=full timeout 900 python -m pytest tests -q -m gpu > $OUT/tcall_6_pytest_auto_full.log 2>&1
timeout 300 python tests/tools/bench_schedule.py --steps 2000 > $OUT/tcall_7_schedule_cfg3.json 2> $OUT/tcall_7_schedule_cfg3.err
if [ "$1" == "ncu" ]; then
  KERNEL=tensor-full timeout 600 bash profiles/run_ncu.sh r2tc
fi
tail -3 $OUT/tcall_6_pytest_auto_full.log $OUT/tcall_0_selftest_mn.log $OUT/tcall_1_capi.log $OUT/tcall_2_pytest.log $OUT/tcall_2b_networks.log $OUT/tcall_3_fwd.log $OUT/tcall_4_bwd.log
tail -c 600 $OUT/tcall_5_bench_auto.json $OUT/tcall_5_bench_full.json $OUT/tcall_5_bench_full_v2.json
