# launch time of k_score against the block width (worker warps per SM) and the shard size
for f in $1; do for w in $2; do
  RVL_SCORE_WARPS=$w python bench.py --features $f --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/w_${f}_$w.json 2>gpurun_out/w_${f}_$w.err
  python - <<EOF2
import json
try:
    d=json.loads(open("gpurun_out/w_${f}_$w.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("F=$f w=$w", "step %.3f ms"%d["ms_per_step"], "k_score launch %.1f us"%(r["launch_ms"]*1e3), "evals/launch %d"%r["per_launch"]["evaluations"], d["check"]["cmi_checksum"])
except Exception as e: print("$f $w", "ERR", e, open("gpurun_out/w_${f}_$w.err").read()[-600:])
EOF2
done; done
