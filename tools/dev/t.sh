python tools/e2e_sweep.py
ORBGPU_TAPER=2 python tools/e2e_sweep.py
ORBGPU_TAPER=4 python tools/e2e_sweep.py
ORBGPU_CHUNK=32 ORBGPU_STREAMS=4 python tools/e2e_sweep.py
ORBGPU_CHUNK=32 ORBGPU_STREAMS=4 ORBGPU_TAPER=3 python tools/e2e_sweep.py
ORBGPU_CHUNK=64 ORBGPU_STREAMS=3 ORBGPU_TAPER=4 python tools/e2e_sweep.py
ORBGPU_CHUNK=24 ORBGPU_STREAMS=4 python tools/e2e_sweep.py
