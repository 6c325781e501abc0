cd orb_slam2_with_comment_b200; cp liborbgpu.so _base.so; cd ..
python tools/stage_times.py 1024 3 256 2>&1 | tail -1
for v in 224_5 224_6; do cp orb_slam2_with_comment_b200/_var_$v.so orb_slam2_with_comment_b200/liborbgpu.so; echo $v; python -m pytest tests/test_gpu_extract.py -m gpu -x -q -k "kitti or end_to_end or parity or golden or shapes" 2>&1 | tail -1; python tools/stage_times.py 1024 3 256 2>&1 | tail -1; done
cp orb_slam2_with_comment_b200/_base.so orb_slam2_with_comment_b200/liborbgpu.so
