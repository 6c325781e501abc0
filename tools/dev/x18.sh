set -x
python -m pytest tests/test_gpu_match.py tests/test_gpu_ref_matcher.py tests/test_gpu_pipeline.py tests/test_cpp_shell.py tests/test_gpu_stereo.py -m gpu -x -q 2>&1 | tail -3
python tools/quick_short_bench.py 5 2>&1 | head -1
for v in g4 g16; do ORBGPU_LIB=$PWD/tools/_build/liborbgpu_$v.so python tools/quick_short_bench.py 5 2>&1 | head -1; done
