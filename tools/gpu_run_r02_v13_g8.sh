set -x
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v13_bench_g8.json 2>gpurun_out/err13.txt; tail -5 gpurun_out/err13.txt
STROTSS_PEER_AR=0 $T bench.py --gpus 8 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v13_bench_g8_ncclar.json 2>>gpurun_out/err13.txt
python - <<'PY'
import json
for f in ['gpurun_out/r02_v13_bench_g8.json','gpurun_out/r02_v13_bench_g8_ncclar.json']:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['rowshard']
        print(f, round(d['value'],1), 'rowshard', round(r['value'],1), r['ms_per_step'], r['parity']['ok'], r['parity']['scalars_max_rel_diff'])
        for k,v in r['phases_ms_per_step_by_rank'].items():
            if 'exch' in k or 'copy' in k or 'ss_' in k: print('   ', k, v)
    except Exception as e: print(f, 'ERR', e)
PY
