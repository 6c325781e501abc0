set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference > gpurun_out/r2_bench3_ref.json 2> gpurun_out/r2_bench3_ref.err
python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
tail -c 600 gpurun_out/r2_bench3.err
