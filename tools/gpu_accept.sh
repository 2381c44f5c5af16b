#!/bin/bash
# Round-2 pass: compute-sanitizer (memcheck / racecheck / synccheck / initcheck) over the case-optimised and generic kernels on
# small grids, then the full-size 100-step acceptance runs (C4 201x101x101, C5 201^3) against the oracle.
mkdir -p gpurun_out
SEL='test_baseline_config_vs_oracle and (fast or ref_order) and (C5_perf_hllc or C4_droplet_weno5z or C2_shockbubble or C1_sod)'
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -q -x -k "$SEL or test_c_program" > gpurun_out/sanitizer_$tool.pytest.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1) | $(tail -1 gpurun_out/sanitizer_$tool.pytest.log)"
done
rm -f gpurun_out/acceptance.log
MFC5_ACCEPTANCE=1 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k test_full_size_100 > gpurun_out/pytest_accept.log 2>&1
tail -3 gpurun_out/pytest_accept.log; cat gpurun_out/acceptance.log
