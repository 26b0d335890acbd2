#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest.log
tail -6 gpurun_out/r02p_pytest.log
