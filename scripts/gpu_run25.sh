mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all25.log 2>&1; echo "== gpu tests exit $?"; tail -n 6 gpurun_out/t_all25.log
timeout 600 python scripts/conv_microbench.py --batch 64 --kinds fprop,dgrad > gpurun_out/micro_v9.log 2>&1; tail -3 gpurun_out/micro_v9.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v10.log 2>&1; echo "== bench exit $?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_v10.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); r=d['roofline']
    print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'clocks',d['clocks'])
    print('families',{k:(round(v['achieved']),round(v['ms_per_step'],2)) for k,v in r['families'].items()},'conv share',round(r['all_conv']['share_of_step'],3))
else: print(open('gpurun_out/bench_v10.log').read()[-2000:])
PY
