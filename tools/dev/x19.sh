set -x
for v in g2 g1; do ORBGPU_LIB=$PWD/tools/_build/liborbgpu_$v.so python -m pytest tests/test_gpu_match.py -m gpu -x -q 2>&1 | tail -1; ORBGPU_LIB=$PWD/tools/_build/liborbgpu_$v.so python tools/quick_short_bench.py 5 2>&1 | head -1; done
