#!/bin/bash
mkdir -p gpurun_out
SWEEP_EXTRA='[{"BGR_TUNE_SUB": 128}, {"BGR_TUNE_SUB": 128, "BGR_TUNE_PASSIVE_EARLY": 0}, {"BGR_TUNE_VEC": 1, "BGR_TUNE_MINB": 8}]' timeout 600 python scripts/sync_sweep.py stress_100k_d8 > gpurun_out/r02q_sweep.jsonl 2> gpurun_out/r02q_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02q_sweep.jsonl | cut -c1-430
