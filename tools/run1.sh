set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(time python bench.py --steps 5 --warmup 3) > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 900 gpurun_out/bench_ref.json
PANTAS_TILE_BYTES=8992 python tools/prof_step.py --pairs 5000000 --steps 3 --ladder > gpurun_out/prof_step.log 2>&1; tail -8 gpurun_out/prof_step.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cli --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
PANTAS_TILE_BYTES=8992 ncu --set full --clock-control none --import-source on -k regex:augment_team -s 1 -c 1 -o gpurun_out/prof_team python tools/prof_step.py --pairs 5000000 --steps 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
