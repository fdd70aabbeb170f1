run() { env "$@" python tools/gpu_phase.py 3584 9472 8 2>&1 | tail -1; }
for rep in 1 2; do
run PHASE_DENT=0 GRPO_L2_HINTS=1
run PHASE_DENT=1 GRPO_L2_HINTS=1
run PHASE_DENT=0 GRPO_L2_HINTS=0
run PHASE_DENT=1 GRPO_L2_HINTS=0
done
nvidia-smi --query-gpu=temperature.gpu,power.draw,clocks.sm --format=csv
