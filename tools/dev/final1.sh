set -x
python bench.py --impl reference > gpurun_out/r2_bench5_ref.json 2> gpurun_out/r2_bench5_ref.err
python bench.py > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
tail -c 300 gpurun_out/r2_bench5.err
python bench.py --steps 2 --warmup 1 > gpurun_out/b2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_v5_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b2_ncu.log 2>&1
tail -c 200 gpurun_out/b2_ncu.log
python tools/quick_bench.py 1024 2 256 > gpurun_out/qb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ -s 36 -c 12 -o gpurun_out/r2_ext_v8_B1024 -f python tools/quick_bench.py 1024 2 256 > gpurun_out/qb_ncu.log 2>&1
cat gpurun_out/qb.log; tail -n 2 gpurun_out/qb_ncu.log
