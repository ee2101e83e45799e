set -x
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T tests/multi_gpu_check.py > gpurun_out/r02_v13_multi_gpu_parity_g2.log 2>&1; tail -3 gpurun_out/r02_v13_multi_gpu_parity_g2.log
$T bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v13_bench_g2.json 2>gpurun_out/err13.txt; tail -5 gpurun_out/err13.txt
STROTSS_PEER_AR=0 $T bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v13_bench_g2_ncclar.json 2>>gpurun_out/err13.txt
python - <<'PY'
import json
for f in ['gpurun_out/r02_v13_bench_g2.json','gpurun_out/r02_v13_bench_g2_ncclar.json']:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['rowshard']
        print(f, round(d['value'],1), 'rowshard', round(r['value'],1), r['ms_per_step'], r['parity']['ok'], r['parity']['scalars_max_rel_diff'], {k:v for k,v in r['phases_ms_per_step'].items() if 'exch' in k or 'copy' in k})
    except Exception as e: print(f, 'ERR', e)
PY
