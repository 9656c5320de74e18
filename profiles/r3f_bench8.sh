# The driver's multi-GPU bench invocation at N = 8 on the final build (independent pairs + the partitioned sub-record).
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r3f_bench_8gpu.json 2> gpurun_out/r3f_bench_8gpu.err; echo "rc $?"; tail -c 300 gpurun_out/r3f_bench_8gpu.err
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/r3f_bench_8gpu.json') if l.startswith('{')]
d=json.loads(lines[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}); print(d.get('partitioned'))
PY
