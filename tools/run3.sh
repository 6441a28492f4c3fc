set -x
mkdir -p gpurun_out
for L in 0 1; do for g in 8192 8193 7168 6144; do PANTAS_LOOSE=$L PANTAS_TEAM_TILE=$g python tools/prof_step.py --pairs 5000000 --steps 4 2>&1 | grep -E "fast kernel" | sed "s/^/loose $L geo $g: /"; done; done
