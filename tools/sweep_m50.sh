python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "m50" 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload m50_50k_basins_x_3650d"
for extra in "" "--maxreg 128" "--maxreg 255" "--block 64" "--block 256"; do
  echo "== $extra"
  $B $extra 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
r=d['roofline']; print('value %.3e frac %.3f kernel_ms %.3f regs %d blocks/SM %d spill %d' % (d['value'], r['frac'], r['kernel_ms'], r['registers'], r['blocks_per_sm'], r['spill_bytes']))"
done
