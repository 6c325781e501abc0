set -x
python -m pytest tests/test_gpu_extract.py tests/test_reference_golden.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
for t in 128 256; do echo "THREADS $t"; ORBGPU_OCT_THREADS=$t python tools/stage_times.py 1024 3 256; done
python tools/quick_bench.py 1024 5 256 2>&1 | tail -2
python tools/quick_bench.py 1 20 1 2>&1 | tail -2
ncu --set full --clock-control none --import-source on --kernel-name regex:k_octree -c 1 -o gpurun_out/r2_oct2_v2 -f python tools/quick_bench.py 1024 1 256 > gpurun_out/run4_ncu.log 2>&1
