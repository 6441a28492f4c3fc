set -x
mkdir -p gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-cli --no-cpu-baseline > gpurun_out/bench_dmfull_n$N.json 2> gpurun_out/bench_dm_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_dmfull_n$N.json').read().strip().split('\n')[-1])
print('N', d['n_gpus'], 'value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'parity', d['parity_checked'], 'kernel ms', d['roofline']['kernel_ms'])
PY
