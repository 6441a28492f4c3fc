set -x
mkdir -p gpurun_out
python tools/prof_step.py --pairs 5000000 --preset gene-panel --seed 1005 --steps 3 2>&1 | grep -E "^pass 2|fast kernel|deferred"
PANTAS_TILE_BYTES=8208 python tools/prof_step.py --pairs 5000000 --preset gene-panel --seed 1005 --steps 3 2>&1 | grep -E "^pass 2|fast kernel|deferred"
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
