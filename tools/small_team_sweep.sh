# K1 time per step of the small network (BASELINE configs[1] shape) for team widths x batch sizes
for E in ${ENVS:-4096 32768 262144}; do
for T in ${TEAMS:-1 8 32}; do
  python bench.py --workload small --envs $E --team $T --no-cpu --no-e2e --no-spot-check --steps 3 --warmup 3 2>/dev/null | tail -n1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('E', $E, 'team', $T, '->', d['config'].get('team_size'), 'value %.1f M' % (d['value']/1e6), 'k1_ms %.4f' % r.get('k1_ms_per_launch'), [round(k['ms_per_launch'],4) for k in r.get('kernels')], d.get('cuda_graph_episode'))
"
done
done
