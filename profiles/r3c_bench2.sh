# The driver's multi-GPU bench invocation at N = 2: independent pairs + the partitioned sub-record.
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r3c_bench_2gpu.json 2> gpurun_out/r3c_bench_2gpu.err; echo "rc $?"; tail -c 400 gpurun_out/r3c_bench_2gpu.err
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/r3c_bench_2gpu.json') if l.startswith('{')]
d=json.loads(lines[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}); print(d.get('partitioned')); print(d['e2e'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | tail -2 | cut -c1-300
