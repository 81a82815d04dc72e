#!/bin/bash
# one compute-sanitizer tool per call (B200_PROFILING.md): memcheck over a small selection of the Detect-path tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_detect_paths_gpu.py -m gpu -x -q -k "four_classes or degenerate or alignment or stage1_alone" > gpurun_out/r2san_plain.log 2>&1
echo "plain rc $?"; tail -2 gpurun_out/r2san_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r2san_memcheck.log python -m pytest tests/test_detect_paths_gpu.py tests/test_tracker_gpu.py tests/test_siblings.py -m gpu -x -q -k "four_classes or degenerate or stage1_alone or (alignment and 750) or (alignment and 1601) or distance_mode or frames_chain or f64_fixture or (beyond and 9000)" > gpurun_out/r2san_pytest.log 2>&1
echo "memcheck rc $?"; tail -3 gpurun_out/r2san_pytest.log; tail -5 gpurun_out/r2san_memcheck.log
