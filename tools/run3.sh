set -x
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 800 python bench.py --workload gene-panel-500M --steps 2 --warmup 3 --no-cli --no-cpu-baseline > gpurun_out/bench_genepanel_500M.json 2> gpurun_out/bench_gp500.err; tail -c 1500 gpurun_out/bench_genepanel_500M.json; tail -5 gpurun_out/bench_gp500.err
