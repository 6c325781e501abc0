set -x
python tools/stage_times.py 1024 3 256 2>&1 | tail -1
for v in r10 r12 d6x32 d8x16 d4x32 d4x16; do ORBGPU_LIB=$PWD/tools/_build/liborbgpu_$v.so python tools/stage_times.py 1024 3 256 2>&1 | tail -1; done
python tools/quick_bench.py 1024 2 256 > gpurun_out/qb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ -s 36 -c 12 -o gpurun_out/r2_ext_v6_B1024 -f python tools/quick_bench.py 1024 2 256 > gpurun_out/qb_ncu.log 2>&1
cat gpurun_out/qb.log; tail -2 gpurun_out/qb_ncu.log
