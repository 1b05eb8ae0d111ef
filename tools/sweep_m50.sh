python -m pytest tests/test_gpu_parity.py -q -x -k "m50 or objective" 2>&1 | tail -2
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload m50_50k_basins_x_3650d"
for extra in "" "--maxreg 128" "--maxreg 255" "--maxreg 255 --block 64" "--maxreg 168 --block 64"; do
  echo "== $extra"
  $B $extra 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
r=d['roofline']; print('value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes']))"
done
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload hbv_100k_basins_x_730d"
for extra in "" "--maxreg 100" "--block 64"; do
  echo "== hbv $extra"
  $B $extra 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
r=d['roofline']; print('value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes']))"
done
