python tools/e2e_sweep.py
for c in 24 32 64 96 128; do for s in 2 3 4 6; do ORBGPU_CHUNK=$c ORBGPU_STREAMS=$s python tools/e2e_sweep.py; done; done
