set -x
mkdir -p gpurun_out
python bench.py --workload dm-full-10M-bq --steps 5 --warmup 3 --no-cli --no-cpu-baseline > gpurun_out/bench_bq.json 2> gpurun_out/bench_bq.err; tail -c 1300 gpurun_out/bench_bq.json; tail -3 gpurun_out/bench_bq.err
