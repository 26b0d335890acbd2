SWEEP_EXTRA='[{"BGR_TUNE_SUB":128},{"BGR_TUNE_SUB":128,"BGR_TUNE_PASSIVE_TMA":0},{"BGR_TUNE_PASSIVE_TMA":0},{"BGR_TUNE_SUB":128,"BGR_TUNE_PASSIVE_TMA":0,"BGR_TUNE_MINB":8},{"BGR_TUNE_SUB":128,"BGR_TUNE_PASSIVE_TMA":0,"BGR_TUNE_STAGGER_NS":0}]' timeout 600 python scripts/sync_sweep.py stress_100k_d8 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if 'error' in d: print(d); continue
    print(d['env'], 'sync %.1f' % d['sync_us_per_tick'], 'kernel %.1f' % d['kernel']['kernel_us_median'], 'pipelined %.1f' % d['pipelined_us_per_tick'])
"
