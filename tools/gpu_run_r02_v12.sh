set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_v12_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_v12_pytest_gpu.log
python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02_v12_bench.json 2>gpurun_out/err12.txt; tail -3 gpurun_out/err12.txt
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_v12_bench.json') if l.startswith('{')][-1]); print(round(d['value'],1), d['phases_ms_per_step'], d['phases_summary'])"
