set -x
python tools/quick_bench.py 1024 2 256 > gpurun_out/qb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ -s 36 -c 12 -o gpurun_out/r2_ext_v11_B1024 -f python tools/quick_bench.py 1024 2 256 > gpurun_out/qb_ncu.log 2>&1
cat gpurun_out/qb.log; tail -n 2 gpurun_out/qb_ncu.log
