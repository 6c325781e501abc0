python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
python tools/multi_bench.py > gpurun_out/r2b_multi_dispatcher.jsonl 2> gpurun_out/multi.err; tail -2 gpurun_out/multi.err
python bench.py --gpus 2 --no-matching --no-vocabulary > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; tail -c 300 gpurun_out/r2b_bench_n2.err
cat gpurun_out/r2b_multi_dispatcher.jsonl | cut -c1-200
cut -c1-400 gpurun_out/r2b_bench_n2.json
