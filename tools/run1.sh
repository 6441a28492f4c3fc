set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cli.py tests/test_gpu_parity.py -x -q -k "cli or device_gfa or bench_scale or gz or stdin or batch" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cli.json 2> gpurun_out/bench_cli.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cli.json').read().strip().split('\n')[-1])
print(d['cli']); print(d['e2e']); print(d['value'], d['ms_per_step'])
PY
tail -3 gpurun_out/bench_cli.err
