set -x
python -m pytest tests/test_gpu_extract.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
python tools/stage_times.py 1024 3 256 2>&1 | tail -1
ORBGPU_LIB=$PWD/tools/_build/liborbgpu_fs1.so python tools/stage_times.py 1024 3 256 2>&1 | tail -1
python tools/quick_bench.py 1024 5 256 2>&1 | head -1
python tools/quick_bench.py 1024 2 256 > gpurun_out/qb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ -s 36 -c 12 -o gpurun_out/r2_ext_v7_B1024 -f python tools/quick_bench.py 1024 2 256 > gpurun_out/qb_ncu.log 2>&1
tail -2 gpurun_out/qb_ncu.log
