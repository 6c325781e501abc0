python -m pytest tests/test_gpu_extract.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
ORBGPU_FAST_CTAS_PER_SM=4 python tools/stage_times.py 1024 3 256 2>&1 | tail -1
