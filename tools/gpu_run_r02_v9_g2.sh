set -x
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $T tests/multi_gpu_check.py > gpurun_out/r02_v9_multi_gpu_parity_g2.log 2>&1; tail -4 gpurun_out/r02_v9_multi_gpu_parity_g2.log
timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v9_bench_g2.json 2>gpurun_out/err9.txt; tail -5 gpurun_out/err9.txt
STROTSS_SHARD_COV=0 timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v9_bench_g2_nocov.json 2>>gpurun_out/err9.txt
python - <<'PY'
import json
for f in ['gpurun_out/r02_v9_bench_g2.json','gpurun_out/r02_v9_bench_g2_nocov.json']:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['rowshard']
        print(f, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'rowshard', round(r['value'],1), r['ms_per_step'], r['parity'], r['phases_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
