set -x
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T tests/multi_gpu_check.py > gpurun_out/r02_v6_multi_gpu_parity.log 2>&1; tail -15 gpurun_out/r02_v6_multi_gpu_parity.log
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v6_bench_g2.json 2>gpurun_out/err6.txt; tail -5 gpurun_out/err6.txt
STROTSS_SHARD_SYM=0 timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v6_bench_g2_rect.json 2>>gpurun_out/err6.txt
python - <<'PY'
import json
for f in ['gpurun_out/r02_v6_bench_g2.json','gpurun_out/r02_v6_bench_g2_rect.json']:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['rowshard']
        print(f, round(d['value'],1), 'rowshard', round(r['value'],1), r['ms_per_step'], r['parity'], r['phases_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
