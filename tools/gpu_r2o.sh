#!/bin/bash
# Round-2 pass O: compute-sanitizer over the code added late in round 2: viscous sweeps (viscous_face, full ghost fill), order 7, templated
# k_cbc, division-light wave_speeds = 2 / low_Mach solve (generic kernels), the x sweep's segment list.
mkdir -p gpurun_out
SEL='(test_viscous_strict_vs_oracle and (8EAC3DA7 or 9CB03CEF or 1C0780C8)) or (test_golden and (2A523AC1 or 345A94C0 or 32D0F235 or 07C1FECC or 1A78D4CC or 66CB22E3 or 4D88A48D)) or (test_order7_strict_vs_oracle) or (test_baseline_config_vs_oracle and fast and (C5_perf_hllc or C2_shockbubble))'
for tool in memcheck initcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer2_$tool.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer2_$tool.pytest.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer2_$tool.log | tail -1) | $(tail -1 gpurun_out/sanitizer2_$tool.pytest.log)"
done
