set -x
G=$1
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
for V in "X=0" "STROTSS_SHARD_SIDE=0"; do
env $V $T bench.py --gpus $G --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v20_bench_g${G}_$V.json 2>gpurun_out/err20.txt; tail -5 gpurun_out/err20.txt
python - <<PY
import json
f='gpurun_out/r02_v20_bench_g${G}_$V.json'
try:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['rowshard']
    print('$V', round(d['value'],1), 'rowshard', round(r['value'],1), r['ms_per_step'], r['parity']['ok'], r['parity']['scalars_max_rel_diff'], r['parity']['own_grad_rows_rel_diff'])
    print('   ', r['phases_ms_per_step'])
except Exception as e: print(f, 'ERR', e)
PY
done
