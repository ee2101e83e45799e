set -x
T="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
time $T bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r02_v17_bench_g2_full.json 2>gpurun_out/err17.txt; tail -5 gpurun_out/err17.txt
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_v17_bench_g2_full.json') if l.startswith('{')][-1]); r=d['rowshard']
print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'rowshard', round(r['value'],1), r['ms_per_step'], r['parity']['ok'], r.get('transport'))
print(d.get('extra'))
print(d.get('cpu_baseline'))
PY
time $T bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_v17_ref_g2.json 2>>gpurun_out/err17.txt; cat gpurun_out/r02_v17_ref_g2.json | cut -c1-600
