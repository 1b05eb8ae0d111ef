set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
for extra in "" "--maxreg 80" "--maxreg 64" "--block 64" "--block 256" "--stages 2" "--stages 8" "--block 64 --maxreg 80" "--block 256 --maxreg 80"; do
  echo "== $extra"
  $B $extra 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
r=d['roofline']; print('value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes']))"
done
