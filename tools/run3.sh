set -x
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --workload hs-chr1-100M-strong --steps 3 --warmup 3 --no-cli --no-cpu-baseline > gpurun_out/bench_hschr1_100M_n$N.json 2> gpurun_out/bench_hs_n$N.err; tail -c 1800 gpurun_out/bench_hschr1_100M_n$N.json; tail -3 gpurun_out/bench_hs_n$N.err
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cli --no-cpu-baseline > gpurun_out/bench_dmfull_n$N.json 2> gpurun_out/bench_dm_n$N.err; tail -c 1800 gpurun_out/bench_dmfull_n$N.json; tail -3 gpurun_out/bench_dm_n$N.err
if [ "$N" = "8" ]; then
$TR bench.py --gpus $N --workload hs-wg-25M --tables --steps 2 --warmup 3 > gpurun_out/bench_hswg_n$N.json 2> gpurun_out/bench_wg_n$N.err; tail -c 1800 gpurun_out/bench_hswg_n$N.json; tail -3 gpurun_out/bench_wg_n$N.err
fi
