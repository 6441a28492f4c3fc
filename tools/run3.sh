set -x
mkdir -p gpurun_out
N=${N:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/prof_reduce.py > gpurun_out/prof_reduce_n$N.log 2>&1
grep -v "^W\|OMP" gpurun_out/prof_reduce_n$N.log | tail -20
