#!/bin/bash
# round 2: compute-sanitizer over the small-shape GPU tests (stepper TMA + fallback paths, wave grid kernel, IRF tiled +
# composition, pack kernels, DE kernels, host-pipelined hb_run).  One log per tool under gpurun_out/.
mkdir -p gpurun_out
SEL='test_tma_and_fallback_paths_agree_bitwise or test_multinode_matches_oracle or test_wave_grid_kernel_matches_oracle_and_generic_path or test_sparse_route_convolution_matches_oracle or test_route_kernel_composition or test_long_destinations or test_pipelined_host_run_is_bit_identical or test_standalone_neural_flux or test_de_kernels or test_ensemble_objective'
for tool in memcheck synccheck initcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_calibration.py tests/test_gpu_rasters.py -m gpu -x -q -k "$SEL or pack or d8" \
    > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r2_sanitizer_summary.txt
  tail -4 gpurun_out/r2_sanitizer_$tool.log
done
