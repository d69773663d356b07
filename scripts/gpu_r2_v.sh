#!/bin/bash
# round-2 visit V: full validation at head -- GPU suite, smoke, driver-form bench (both arms), ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2v_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2v_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2v_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2v_smoke.log
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench exit $?"; cut -c1-330 gpurun_out/r2v_bench.json
GRAPH=0 WARM=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2v_launches.csv python scripts/profile_step.py > gpurun_out/r2v_profile_step.log 2>&1; echo "ncu exit $?"; tail -1 gpurun_out/r2v_profile_step.log
python scripts/summarize_launches.py gpurun_out/r2v_launches.csv > gpurun_out/r2v_launches_summary.txt 2>&1; head -50 gpurun_out/r2v_launches_summary.txt
gzip -f gpurun_out/r2v_launches.csv
