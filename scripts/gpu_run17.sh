mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/t_all17.log 2>&1; echo "== gpu tests exit $?"; tail -n 4 gpurun_out/t_all17.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v7.log 2>&1; echo "== bench exit $?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_v7.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); r=d['roofline']
    print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'clocks',d['clocks'],'cpu',d['cpu_baseline'])
    print('families',{k:(round(v['achieved']),round(v['ms_per_step'],2)) for k,v in r['families'].items()},'conv share',round(r['all_conv']['share_of_step'],3))
PY
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 205 -c 215 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"
