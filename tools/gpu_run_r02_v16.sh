set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "self_similarity or total_against or full_size_against" > gpurun_out/r02_v16_pytest_subset.log 2>&1; tail -3 gpurun_out/r02_v16_pytest_subset.log
B="python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline"
$B > gpurun_out/r02_v16_bench.json 2>gpurun_out/err16.txt
STROTSS_P_STREAM=1 $B > gpurun_out/r02_v16_bench_pstream.json 2>>gpurun_out/err16.txt
$B > gpurun_out/r02_v16_bench_b.json 2>>gpurun_out/err16.txt
STROTSS_P_STREAM=1 $B > gpurun_out/r02_v16_bench_pstream_b.json 2>>gpurun_out/err16.txt
tail -3 gpurun_out/err16.txt
for f in gpurun_out/r02_v16_bench*.json; do python -c "
import json; d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); p=d['phases_ms_per_step']; print('$f', round(d['value'],1), 'ss1', p['ss_stage1_gemm'], 'ss2', p['ss_stage2_gemm'], 'ss_misc', p['ss_misc'], d['phases_summary'])"; done
