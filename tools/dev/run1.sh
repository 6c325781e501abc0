set -x
python -m pytest tests/test_gpu_extract.py tests/test_reference_golden.py -m gpu -x -q 2>&1 | tail -3
for f in 0 4 5; do echo "FUSE $f"; ORBGPU_FUSE_OCT=$f python tools/quick_bench.py 1024 5 256 2>&1 | tail -2; done
ORBGPU_FUSE_OCT=0 python tools/stage_times.py 1024 3 256
ORBGPU_FUSE_OCT=4 python tools/stage_times.py 1024 3 256
