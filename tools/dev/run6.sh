set -x
ORBGPU_DEBUG=1 python -m pytest tests/test_gpu_extract.py tests/test_reference_golden.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -8
for r in 6 8 12 16; do echo "BAND ROWS $r"; ORBGPU_DEBUG=1 ORBGPU_PYR_BAND_ROWS=$r python tools/stage_times.py 1024 3 256 2>&1 | tail -2; done
ORBGPU_PYR_FUSED=0 python tools/stage_times.py 1024 3 256
python tools/quick_bench.py 1024 5 256 2>&1 | tail -2
python tools/quick_bench.py 1 20 1 2>&1 | tail -2
