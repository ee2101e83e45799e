set -x
# last GPU-minutes of the round: as much of `pytest -m gpu` on the final build as fits (SIGINT so that pytest still prints its summary)
timeout -s INT 130 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_full_size_against_the_fp64_restatement --deselect tests/test_gpu_parity.py::test_alternative_kernel_paths > gpurun_out/r02_v22_pytest_gpu_partial.log 2>&1
tail -5 gpurun_out/r02_v22_pytest_gpu_partial.log
