set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_step.py --pairs 5000000 --steps 3 --ladder > gpurun_out/prof_step.log 2>&1; tail -9 gpurun_out/prof_step.log
PANTAS_LOOSE=1 PANTAS_TEAM_TILE=8193 python tools/prof_step.py --pairs 5000000 --steps 3 2>&1 | grep "fast kernel" | sed "s/^/loose 8193: /"
ncu --set full --clock-control none --import-source on -k regex:augment_team -s 1 -c 1 -o gpurun_out/prof_team python tools/prof_step.py --pairs 5000000 --steps 2 > gpurun_out/ncu_full.log 2>&1
