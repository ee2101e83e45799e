set -x
# the last GPU-seconds: the 19 alternative-kernel-path settings, six subprocess tests at a time (pytest-xdist)
timeout -s INT 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -n 6 -k "alternative_kernel_paths" > gpurun_out/r02_v22_pytest_gpu_alt_paths.log 2>&1
tail -6 gpurun_out/r02_v22_pytest_gpu_alt_paths.log
