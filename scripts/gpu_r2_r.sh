#!/bin/bash
# round-2 visit R: ncu launch list of one GRPO group at config 2 (shares of the step), new GPU tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 600 -x -k "stat_tracker or advantage" > gpurun_out/r2r_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2r_tests.log
GRAPH=0 WARM=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2r_launches.csv python scripts/profile_step.py > gpurun_out/r2r_profile_step.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/r2r_profile_step.log
python scripts/summarize_launches.py gpurun_out/r2r_launches.csv > gpurun_out/r2r_launches_summary.txt 2>&1; head -45 gpurun_out/r2r_launches_summary.txt
gzip -f gpurun_out/r2r_launches.csv
