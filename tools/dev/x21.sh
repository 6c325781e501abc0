set -x
ORBGPU_LIB=$PWD/tools/_build/liborbgpu_dsw.so timeout 300 python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -2
ORBGPU_LIB=$PWD/tools/_build/liborbgpu_dsw.so timeout 120 python tools/stage_times.py 1024 3 256 2>&1 | tail -1
python tools/stage_times.py 1024 3 256 2>&1 | tail -1
