set -x
G=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $T bench.py --gpus $G --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v11_bench_g$G.json 2>gpurun_out/err11.txt; tail -5 gpurun_out/err11.txt
python - <<PY
import json
for f in ['gpurun_out/r02_v11_bench_g$G.json']:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['rowshard']
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'rowshard', round(r['value'],1), r['ms_per_step'], r['parity'], r['phases_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
