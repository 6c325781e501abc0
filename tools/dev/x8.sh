set -x
ncu --set full --clock-control none --import-source on -k regex:k_blur_tma -s 3 -c 1 -o gpurun_out/r2_blur_tile_rolled -f python tools/quick_bench.py 1024 1 256 > gpurun_out/qb_ncu1.log 2>&1
ORBGPU_LIB=$PWD/tools/_build/liborbgpu_band.so ncu --set full --clock-control none --import-source on -k regex:k_blur_tma -s 3 -c 1 -o gpurun_out/r2_blur_band_rolled -f python tools/quick_bench.py 1024 1 256 > gpurun_out/qb_ncu2.log 2>&1
tail -2 gpurun_out/qb_ncu1.log gpurun_out/qb_ncu2.log
