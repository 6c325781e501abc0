set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ORBGPU_RESIZE_TMA=0 python bench.py --no-matching --no-vocabulary > gpurun_out/b_tma0.json 2> gpurun_out/b_tma0.err
ORBGPU_RESIZE_TMA=1 python bench.py --no-matching --no-vocabulary > gpurun_out/b_tma1.json 2> gpurun_out/b_tma1.err
python - <<'PY'
import json
for f in ("gpurun_out/b_tma0.json","gpurun_out/b_tma1.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["stage_ms"])
    except Exception as e: print(f, "ERR", e)
PY
