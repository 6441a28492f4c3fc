set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
