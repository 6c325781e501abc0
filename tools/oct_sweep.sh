for thr in 128 256; do for sm in 0 32768 49152 65536 98304; do echo "threads $thr smem $sm"; ORBGPU_OCT_THREADS=$thr ORBGPU_OCT_SMEM=$sm python tools/stage_times.py 1024 3; done; done
