for v in "$@"; do
  if [ "$v" = "tree" ]; then unset RVL_LIB; else export RVL_LIB=variants/$v.so; fi
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/ab_$v.json 2>gpurun_out/ab_$v.err
  python - <<EOF2
import json
try:
    d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
    print("$v", "%.1fM"%(d["value"]/1e6), "step %.3f ms"%d["ms_per_step"], "launch %.4f ms"%d["roofline"]["launch_ms"], d["check"]["cmi_checksum"])
except Exception as e: print("$v", "ERR", e, open("gpurun_out/ab_$v.err").read()[-600:])
EOF2
done
