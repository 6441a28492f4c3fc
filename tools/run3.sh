set -x
mkdir -p gpurun_out
N=${N:-8}
cat /sys/devices/system/node/online; nvidia-smi topo -m 2>&1 | head -14
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
PANTAS_NUMA=0 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cli --no-cpu-baseline > gpurun_out/bench_n${N}_numa0.json 2> gpurun_out/bench_n${N}_numa0.err
$TR bench.py --gpus $N --steps 3 --warmup 3 --no-cli --no-cpu-baseline > gpurun_out/bench_n${N}_numa1.json 2> gpurun_out/bench_n${N}_numa1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_n*_numa?.json')):
    d=json.loads(open(f).read().strip().split('\n')[-1])
    print(f, 'value %.3g'%d['value'], 'ms %.2f'%d['ms_per_step'], 'e2e %.3g'%d['e2e']['value'], 'e2e ms %.1f'%d['e2e']['ms_per_step'], 'h2d GB/s %.0f'%d['e2e']['h2d_gb_per_s_aggregate'], d.get('numa'))
PY
