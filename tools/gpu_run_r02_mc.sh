set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o /tmp/mc_probe tools/mc_probe.cu -lcuda || exit 1
timeout 120 /tmp/mc_probe > gpurun_out/r02_mc_probe.jsonl 2>&1; cat gpurun_out/r02_mc_probe.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum --clock-control none --csv --log-file gpurun_out/r02_mc_probe_ncu.csv /tmp/mc_probe > /dev/null 2>&1
tail -3 gpurun_out/r02_mc_probe_ncu.csv | cut -c1-300
