python -m pytest tests/test_gpu_extract.py tests/test_reference_golden.py -m gpu -x -q 2>&1 | tail -3
python tools/stage_times.py 1024 3 256 2>&1 | tail -1
