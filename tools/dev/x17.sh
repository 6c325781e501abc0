set -x
python -m pytest tests/test_gpu_match.py tests/test_gpu_ref_matcher.py tests/test_gpu_pipeline.py tests/test_cpp_shell.py -m gpu -x -q 2>&1 | tail -3
python tools/quick_short_bench.py 5 2>&1 | head -2
ORBGPU_SBP_WARP=1 python tools/quick_short_bench.py 5 2>&1 | head -2
