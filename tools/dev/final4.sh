set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
tail -c 200 gpurun_out/r2_bench9.err
python bench.py --steps 2 --warmup 1 > gpurun_out/b2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_v8_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b2_ncu.log 2>&1
tail -c 200 gpurun_out/b2_ncu.log
python tools/quick_bench.py 1024 2 256 > gpurun_out/qb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_ -s 36 -c 12 -o gpurun_out/r2_ext_v10_B1024 -f python tools/quick_bench.py 1024 2 256 > gpurun_out/qb_ncu.log 2>&1
cat gpurun_out/qb.log; tail -n 2 gpurun_out/qb_ncu.log
python -c "import __graft_entry__ as g; g.smoke()"
