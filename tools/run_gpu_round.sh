set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python tools/sweep.py ${SWEEP:-32768 24576} > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
(time python bench.py --steps 5 --warmup 3) > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1; cat gpurun_out/e2e_breakdown.log
python tools/prof_step.py --pairs 1000000 --steps 3 > gpurun_out/prof_step.log 2>&1; tail -5 gpurun_out/prof_step.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --pairs 1000000 --steps 2 > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:augment_fast -s 1 -c 1 -o gpurun_out/prof_fast python tools/prof_step.py --pairs 1000000 --steps 2 > gpurun_out/ncu_full.log 2>&1
if [ -n "$TRAFFIC" ]; then ncu --set full --clock-control none -k regex:augment_fast -s 1 -c 1 -o gpurun_out/prof_fast_10M python tools/prof_step.py --pairs 5000000 --steps 2 > gpurun_out/ncu_full_10M.log 2>&1; fi
ls -la gpurun_out
