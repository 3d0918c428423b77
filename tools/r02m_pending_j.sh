out=gpurun_out/r02m; mkdir -p $out
( time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_cpp_api_gpu.py -m gpu -q -x ) > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log | cut -c1-300
TAG=r02m bash tools/c1_sample.sh
