#!/bin/bash
# end-of-round evidence on one box: verification (suite, smoke, bench, reference arm), sanitizer, ncu of the generic kernels
bash scripts/gpu_verify.sh
bash scripts/gpu_sanitizer.sh
bash scripts/gpu_generic_ncu.sh
