set -x
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cli --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 1200 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
