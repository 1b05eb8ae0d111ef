nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 20 > gpurun_out/clocks_probe.csv &
SMI=$!
sleep 0.5
python bench.py --steps 60 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print([round(x,2) for x in d['kernel_ms_each']]); print(d['clocks'])"
sleep 0.3
kill $SMI
awk -F, 'NR>1 && $4+0 > 300 {print $2, $3, $4, $5, $6, $7}' gpurun_out/clocks_probe.csv | sort | uniq -c | sort -rn | head -12
