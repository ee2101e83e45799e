set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_v3_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02_v3_pytest_gpu.log
B="python bench.py --steps 40 --warmup 5 --no-extra --no-cpu-baseline"
$B > gpurun_out/r02_v3_bench.json 2>gpurun_out/err.txt
STROTSS_PANEL=2048 $B > gpurun_out/r02_v3_bench_panel2048.json 2>>gpurun_out/err.txt
STROTSS_PANEL=2048 STROTSS_P_PERSIST=1 $B > gpurun_out/r02_v3_bench_panel2048_persist.json 2>>gpurun_out/err.txt
STROTSS_PANEL=3072 STROTSS_P_PERSIST=1 $B > gpurun_out/r02_v3_bench_panel3072_persist.json 2>>gpurun_out/err.txt
STROTSS_P_PERSIST=1 $B > gpurun_out/r02_v3_bench_panel4096_persist.json 2>>gpurun_out/err.txt
tail -5 gpurun_out/err.txt
for f in gpurun_out/r02_v3_bench*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value'],1), d['phases_ms_per_step']['ss_stage1_gemm'], d['phases_ms_per_step']['ss_stage2_gemm'], d['phases_summary'])"; done
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none"
B1="python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline"
ncu $M -k regex:"ss1_pair_merged|gemm2w" -s 11 -c 11 --csv --log-file gpurun_out/r02_v3_dram_panel4096.csv $B1 > /dev/null 2>&1
STROTSS_PANEL=2048 ncu $M -k regex:"ss1_pair_merged|gemm2w" -s 23 -c 23 --csv --log-file gpurun_out/r02_v3_dram_panel2048.csv $B1 > /dev/null 2>&1
STROTSS_PANEL=2048 STROTSS_P_PERSIST=1 ncu $M -k regex:"ss1_pair_merged|gemm2w" -s 23 -c 23 --csv --log-file gpurun_out/r02_v3_dram_panel2048_persist.csv $B1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rows_stats3|rows_emit3|finalize_grad|pal_min2" -s 4 -c 4 -o gpurun_out/r02_v3_hbm_kernels $B1 > /dev/null 2>&1
ls -la gpurun_out/ | tail -12
