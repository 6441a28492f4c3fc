set -x
for m in 1 17 18 20 19; do
  echo "== PANTAS_LOOSE=$m"; PANTAS_LOOSE=$m PANTAS_TILE_BYTES=8992 python tools/prof_step.py --pairs 5000000 --steps 4 2>&1 | grep -E "fast kernel"
done
