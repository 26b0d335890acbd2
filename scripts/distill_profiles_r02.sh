#!/bin/bash
# CPU box, after scripts/gpu_profiles_r02.sh came back: distil gpurun_out/ (scratch) into the tracked profiles/ files.
set -u
cd "$(dirname "$0")/.."
for f in gpurun_out/r02_bench_*.json; do
  [ -s "$f" ] && python - "$f" <<'PY'
import json, sys
p = sys.argv[1]
for line in open(p):
    if line.startswith("{"):
        d = json.loads(line)
        out = "profiles/" + p.split("/")[-1]
        json.dump(d, open(out, "w"), indent=1)
        print("wrote", out)
PY
done
[ -s gpurun_out/r02_timeline.csv ] && cp gpurun_out/r02_timeline.csv profiles/r02_timeline.csv
[ -s gpurun_out/r02_launches.csv ] && grep -v "^==" gpurun_out/r02_launches.csv > profiles/r02_launches.csv
[ -s gpurun_out/r02_generic_world.json ] && cp gpurun_out/r02_generic_world.json profiles/r02_generic_world.json
[ -s gpurun_out/r02_prof_fused.ncu-rep ] && python tools/ncu_summary.py gpurun_out/r02_prof_fused.ncu-rep profiles/r02_k_particles_program stress_1m_d8
[ -s gpurun_out/r02_prof_fused.ncu-rep ] && python tools/ncu_timeline.py gpurun_out/r02_prof_fused.ncu-rep profiles/r02_k_particles_program.timeline.csv 0
[ -s gpurun_out/r02_prof_fused_100k.ncu-rep ] && python tools/ncu_summary.py gpurun_out/r02_prof_fused_100k.ncu-rep profiles/r02_k_particles_program_100k stress_100k_d8
[ -s gpurun_out/r02_prof_tma.ncu-rep ] && python tools/ncu_summary.py gpurun_out/r02_prof_tma.ncu-rep profiles/r02_k_image_tma
[ -s gpurun_out/r02_prof_generic.ncu-rep ] && python tools/ncu_summary.py gpurun_out/r02_prof_generic.ncu-rep profiles/r02_k_generic_program
ls -la profiles | tail -30
