# bench line essentials + launch list summary
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'fwd_ms', d['roofline']['kernel_ms'], 'combine_ms', d['roofline']['combine_ms'], 'e2e_ms', d['e2e']['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph > /dev/null 2>&1
python profiles/summarise_launches.py gpurun_out/launches.csv flow_photo
