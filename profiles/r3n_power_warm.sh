# Warm-started power iterations for the damping factors: 3 steps per system (default) against 2 and 1.
mkdir -p gpurun_out
for cfg in "w3:" "w2:MOF_MG_POWER_ITS_WARM=2" "w1:MOF_MG_POWER_ITS_WARM=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 4 --warmup 3 --quick > gpurun_out/r3n_bench_$name.json 2> gpurun_out/r3n_bench_$name.err; echo "== $name rc $?"
  python -c "
import json; d=json.load(open('gpurun_out/r3n_bench_$name.json')); print(round(d['ms_per_step'],1), d['pcg']['flow_iterations_per_alignment'], d['pcg']['smooth_iterations_per_alignment'])"
done
