set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ORBGPU_RESIZE_PP=1 python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -2
python tools/stage_times.py 1024 3 256 2>&1 | tail -1
ORBGPU_RESIZE_PP=1 python tools/stage_times.py 1024 3 256 2>&1 | tail -1
python tools/quick_bench.py 1024 5 256 2>&1 | head -1
ORBGPU_RESIZE_PP=1 python tools/quick_bench.py 1024 5 256 2>&1 | head -1
for pad in 3072 8192 24000 40000; do ORBGPU_RESIZE_PP=1 ORBGPU_OCT_PAD=$pad python tools/quick_bench.py 1024 5 256 2>&1 | head -1; done
