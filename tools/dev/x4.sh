set -x
python -m pytest tests/test_gpu_extract.py tests/test_gpu_pipeline.py tests/test_gpu_stereo.py -m gpu -x -q 2>&1 | tail -3
python tools/stage_times.py 1024 3 256 2>&1 | tail -1
python tools/quick_bench.py 1024 5 256 2>&1 | head -1
python tools/quick_bench.py 64 20 64 2>&1 | head -1
python tools/quick_bench.py 1 50 1 2>&1 | head -1
