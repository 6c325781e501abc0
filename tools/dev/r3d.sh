export ORBGPU_FRAMES_CACHE=/tmp/frames256.npy
python tools/stage_times.py 1024 5 256 2>&1 | tail -1
echo box16; ORBGPU_LIB=tools/_build/liborbgpu_box16.so python tools/stage_times.py 1024 5 256 2>&1 | tail -1
python tools/stage_times.py 1024 5 256 2>&1 | tail -1
echo box16; ORBGPU_LIB=tools/_build/liborbgpu_box16.so python tools/stage_times.py 1024 5 256 2>&1 | tail -1
python tools/parity_report.py 9 2>&1 | cut -c1-330
