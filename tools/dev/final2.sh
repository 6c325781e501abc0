set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
tail -c 200 gpurun_out/r2_bench6.err
python tools/bench_configs.py --scale 1.0 > gpurun_out/r2_configs_1gpu_full_v2.jsonl 2> gpurun_out/configs.err
tail -c 300 gpurun_out/configs.err
python tools/quick_short_bench.py 3 > gpurun_out/qs.log 2>&1 && ncu --set full --clock-control none -k regex:"k_sbp|k_grid|k_tri|k_frustum|k_stereo" -c 24 -o gpurun_out/r2_shortlist_v2 -f python tools/quick_short_bench.py 1 > gpurun_out/qs_ncu.log 2>&1
cat gpurun_out/qs.log
python tools/quick_match_bench.py 4096 256 > gpurun_out/qm.log 2>&1 && ncu --set full --clock-control none -k regex:"k_bow" -c 8 -o gpurun_out/r2_match_v2_P4096 -f python tools/quick_match_bench.py 4096 256 > gpurun_out/qm_ncu.log 2>&1
head -1 gpurun_out/qm.log
