# A/B of one environment switch on the default bench: tools/ab_env.sh VAR [extra bench args]
v=$1; shift
for e in 0 1 0 1; do
  if [ $e = 1 ]; then export $v=1; else unset $v; fi
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v=$e', '%.1fM'%(d['value']/1e6), 'step %.4f ms'%d['ms_per_step'], 'k_score %.1f us'%(d['roofline']['launch_ms']*1e3), d['check']['cmi_checksum'], 'exp/launch %.0f' % d['roofline']['per_launch']['exp'])"
done
