set -x
python bench.py > gpurun_out/r02_final_bench.json 2>gpurun_out/err_final.txt; tail -3 gpurun_out/err_final.txt
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02_final_bench.json') if l.startswith('{')][-1]); print(round(d['value'],1), round(d['e2e']['value'],1), d['phases_summary'], d['roofline']['frac'], d['roofline'].get('traffic'), d.get('extra'), d['cpu_baseline']['value'], d['clocks'])"
B1="python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv $B1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ss1_pair_merged" -s 4 -c 4 -o gpurun_out/r02_final_ss1 python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm2s_kernel|gemm2w_kernel|pal_min2|finalize_grad|rows_emit3|rows_stats3" -s 8 -c 12 -o gpurun_out/r02_final_others python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
