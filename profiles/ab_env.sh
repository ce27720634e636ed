# A/B of an environment switch on the per-kernel eager timings of the training step (one GPU, under gpurun):
#   bash profiles/ab_env.sh VAR "v1 v2 v3" [extra bench.py flags]
# prints, per value, the graph-replayed ms/step and the eager per-kernel milliseconds of the hot entry points
VAR=$1; VALS=$2; shift 2
for v in $VALS; do
  env $VAR=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --long-steps 0 "$@" > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.log
  python - "$VAR" "$v" <<'PY'
import json, sys
var, v = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/ab_{var}_{v}.json"))
    k = d["kernels"]
    names = ["ngp_grid_encode_backward", "ngp_field_sigma_forward", "ngp_field_color_backward_ex", "ngp_ffmlp_backward_ex", "ngp_field_color_forward", "ngp_march_rays_train"]
    print(f"{var}={v}: step {d['ms_per_step']:.3f} ms | " + " ".join(f"{n.replace('ngp_','')}={k[n]['ms_per_step']:.3f}" for n in names if n in k))
except Exception as e:
    print(f"{var}={v}: failed ({e})")
PY
done
