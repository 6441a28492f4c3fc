set -x
mkdir -p gpurun_out
for t in 8192 8960 9216; do
  echo "== dm-full PANTAS_TILE_BYTES=$t"; PANTAS_TILE_BYTES=$t python tools/prof_step.py --pairs 5000000 --steps 4 2>&1 | grep -E "^pass 3|fast kernel"
done
for t in 7680 8192 8448 8704 9216; do
  echo "== gene-panel PANTAS_TILE_BYTES=$t"; PANTAS_TILE_BYTES=$t python tools/prof_step.py --pairs 5000000 --preset gene-panel --seed 1005 --steps 4 2>&1 | grep -E "^pass 3|fast kernel"
done
python bench.py --steps 5 --warmup 3 --no-cli --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 500 gpurun_out/bench_quick.json; head -c 300 gpurun_out/bench_quick.json
