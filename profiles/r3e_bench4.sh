# The driver's multi-GPU bench invocation at N = 4 (independent pairs + the partitioned sub-record), and the GPU suite's texture / command-line tests on the new product link.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_texprep.py tests/test_gpu_cli.py tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r3e_bench_4gpu.json 2> gpurun_out/r3e_bench_4gpu.err; echo "rc $?"; tail -c 300 gpurun_out/r3e_bench_4gpu.err
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/r3e_bench_4gpu.json') if l.startswith('{')]
d=json.loads(lines[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}); print(d.get('partitioned'))
PY
