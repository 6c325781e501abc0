set -x
export ORBGPU_FRAMES_CACHE=/tmp/frames256.npy
python tools/stage_times.py 1024 5 256 2>&1 | tail -1
python tools/stage_times.py 1024 5 256 2>&1 | tail -1
python tools/parity_report.py 24 > gpurun_out/parity24.jsonl 2>&1; cut -c1-330 gpurun_out/parity24.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
tail -c 200 gpurun_out/r2_bench11.err
python -c "import __graft_entry__ as g; g.smoke()"
