mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -q -k "graph or lightning" --tb=short -p no:cacheprovider > gpurun_out/t_graph.log 2>&1; echo "== graph test exit $?"; tail -n 12 gpurun_out/t_graph.log
for mode in "" "--no-graph"; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $mode > gpurun_out/bench_g.log 2>&1; echo "== bench $mode exit $?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_g.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); r=d['roofline']
    print('graph',d.get('cuda_graph'),'value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'clocks',d['clocks'],'launches',d['gpu_launches'])
    print('families',{k:(round(v['achieved']),round(v['ms_per_step'],2)) for k,v in r['families'].items()},'conv share',round(r['all_conv']['share_of_step'],3))
else:
    print(open('gpurun_out/bench_g.log').read()[-1500:])
PY
done
