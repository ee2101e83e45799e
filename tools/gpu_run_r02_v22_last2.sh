set -x
# the remaining GPU-seconds: the full-size tests the first call left out (SIGINT so that pytest still prints its summary)
timeout -s INT 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider --durations=8 -k "full_size" > gpurun_out/r02_v22_pytest_gpu_full_size.log 2>&1
tail -16 gpurun_out/r02_v22_pytest_gpu_full_size.log
