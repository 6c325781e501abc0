set -x
ncu --set full --clock-control none --import-source on --kernel-name regex:k_octree -c 2 -o gpurun_out/r2_oct2_v1 -f python tools/quick_bench.py 1024 1 256 > gpurun_out/run3_ncu.log 2>&1
