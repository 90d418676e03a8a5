# launch time of k_score against the shard size (1 GPU): the fixed cost per launch that limits multi-GPU scaling
for f in "$@"; do
  python bench.py --features $f --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/f_$f.json 2>gpurun_out/f_$f.err
  python - <<EOF2
import json
try:
    d=json.loads(open("gpurun_out/f_$f.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("F=$f", "step %.3f ms"%d["ms_per_step"], "k_score launch %.1f us"%(r["launch_ms"]*1e3), "evals/launch %d"%r["per_launch"]["evaluations"], "share %.3f"%r["share_of_step"])
except Exception as e: print("$f", "ERR", e, open("gpurun_out/f_$f.err").read()[-600:])
EOF2
done
