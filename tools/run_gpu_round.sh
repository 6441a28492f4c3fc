# one gpurun call: parity tests, ablation ladder, bench, launch list, one ncu --set full capture of the fast kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python tools/prof_step.py --pairs 1000000 --steps 3 --ladder > gpurun_out/prof_step.log 2>&1; tail -12 gpurun_out/prof_step.log
if [ -z "$NOBENCH" ]; then (time python bench.py --steps 5 --warmup 3) > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err; fi
if [ -z "$NONCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --pairs 1000000 --steps 2 > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:augment_team -s 1 -c 1 -o gpurun_out/prof_team python tools/prof_step.py --pairs 1000000 --steps 2 > gpurun_out/ncu_full.log 2>&1
fi
if [ -n "$TRAFFIC" ]; then ncu --set full --clock-control none -k regex:augment_team -s 1 -c 1 -o gpurun_out/prof_team_10M python tools/prof_step.py --pairs 5000000 --steps 2 > gpurun_out/ncu_full_10M.log 2>&1; fi
ls -la gpurun_out
