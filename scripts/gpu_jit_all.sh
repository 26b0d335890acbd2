bash scripts/gpu_jit.sh
bash scripts/gpu_jit_particles.sh
