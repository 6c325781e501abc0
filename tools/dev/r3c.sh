export ORBGPU_FRAMES_CACHE=/tmp/frames256.npy
python tools/stage_times.py 1024 5 256 2>&1 | tail -1
for v in k32 w8m4 w4m9 k8; do echo $v; ORBGPU_LIB=tools/_build/liborbgpu_$v.so python tools/stage_times.py 1024 5 256 2>&1 | tail -1; done
