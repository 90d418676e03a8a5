for f in 100000 12500; do for e in 0 1 0 1; do
  export RVL_PDL=$e
  python bench.py --features $f --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>gpurun_out/pdl.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('F=$f RVL_PDL=$e', '%.1fM'%(d['value']/1e6), 'step %.4f ms'%d['ms_per_step'], d['check']['cmi_checksum'])" || tail -3 gpurun_out/pdl.err
done; done
