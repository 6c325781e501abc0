set -x
ORBGPU_RESIZE_TMA=2 python -m pytest tests/test_gpu_extract.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -2
for v in 0 1 2; do ORBGPU_RESIZE_TMA=$v python tools/stage_times.py 1024 3 256 2>&1 | tail -1; done
ORBGPU_RESIZE_TMA=2 python tools/quick_bench.py 1024 5 256 2>&1 | head -1
