set -x
export ORBGPU_FRAMES_CACHE=/tmp/frames256.npy
python tools/stage_times.py 1024 5 256 2>&1 | tail -1
for v in p1m4 p1m5 p2m5 p2m6; do ORBGPU_LIB=tools/_build/liborbgpu_$v.so python tools/stage_times.py 1024 5 256 2>&1 | tail -1; done
python tools/stage_times.py 1024 5 256 2>&1 | tail -1
python tools/parity_report.py 48 > gpurun_out/r2_parity_report.jsonl 2> gpurun_out/parity.err; cat gpurun_out/r2_parity_report.jsonl; tail -3 gpurun_out/parity.err
for v in p1m5 p2m6; do ORBGPU_LIB=tools/_build/liborbgpu_$v.so python tools/parity_report.py 9 2>&1 | cut -c1-400; done
python tools/quick_voc_bench.py 1024 3 > gpurun_out/qv.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_voc -c 10 -o gpurun_out/r2_vocab_F1024 -f python tools/quick_voc_bench.py 1024 1 > gpurun_out/qv_ncu.log 2>&1
cat gpurun_out/qv.log; tail -n 2 gpurun_out/qv_ncu.log
