set -x
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
for V in "X=0" "STROTSS_SHARD_REMD_SIDE=1"; do
env $V $T bench.py --gpus 8 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v19_bench_g8_$V.json 2>gpurun_out/err19.txt; tail -5 gpurun_out/err19.txt
python - <<PY
import json
f='gpurun_out/r02_v19_bench_g8_$V.json'
try:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['rowshard']
    print('$V', round(d['value'],1), 'rowshard', round(r['value'],1), r['ms_per_step'], r['parity']['ok'], r['parity']['scalars_max_rel_diff'], r['parity']['own_grad_rows_rel_diff'])
    for k,v in r['phases_ms_per_step_by_rank'].items():
        if 'exch' in k or 'ss_' in k or 'remd' in k: print('   ', k, v)
except Exception as e: print(f, 'ERR', e)
PY
done
