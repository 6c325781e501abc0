ncu --set full --clock-control none --import-source on --kernel-name regex:k_fast_seg -c 1 -o gpurun_out/r2_fast_persist -f python tools/quick_bench.py 1024 1 256 > gpurun_out/p_ncu.log 2>&1
