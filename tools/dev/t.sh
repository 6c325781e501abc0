python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-matching --no-vocabulary > gpurun_out/b5.json 2> gpurun_out/b5.err; tail -c 300 gpurun_out/b5.err
