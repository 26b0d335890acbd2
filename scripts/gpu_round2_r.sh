#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fuzz_requests.py tests/test_gpu_hierarchy.py -m gpu -q -x > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -25 gpurun_out/r02r_pytest.log
