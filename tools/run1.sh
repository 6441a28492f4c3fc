set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "device_gfa or bench_scale" 2>&1 | tail -3
