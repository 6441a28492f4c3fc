set -x
mkdir -p gpurun_out
(time python bench.py --steps 5 --warmup 3) > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 2500 gpurun_out/bench_ref.json; tail -4 gpurun_out/bench_ref.err
