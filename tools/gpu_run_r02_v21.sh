set -x
B="python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline"
for V in "X=0" "STROTSS_SIDE=1" "X=1" "STROTSS_SIDE=1"; do
env $V $B > gpurun_out/r02_v21_bench_$V.json 2>gpurun_out/err21.txt; tail -2 gpurun_out/err21.txt
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_v21_bench_$V.json') if l.startswith('{')][-1]); print('$V', round(d['value'],1), round(d['e2e']['value'],1), d['loss'], d['clocks']['sm_mhz'])"
done
STROTSS_SIDE=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "total_against or full_size_against or alternative" 2>&1 | tail -2
