set -x
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for V in "STROTSS_PEER_WINDOW=0" "STROTSS_SHARD_SYM=0" "STROTSS_SHARD_COV=0" "STROTSS_SHARD_SS1_STREAMS=1"; do
  env $V $T bench.py --gpus 2 --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_v18_g2_$V.json 2>gpurun_out/err18.txt; tail -3 gpurun_out/err18.txt
  python - <<PY
import json
f='gpurun_out/r02_v18_g2_$V.json'
try:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d.get('rowshard', d)
    print('$V', round(r['value'],1), r.get('ms_per_step'), r.get('parity'), r.get('transport'))
except Exception as e: print(f, 'ERR', e)
PY
done
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
