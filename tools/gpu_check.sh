python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/$1_tests.log; python bench.py > gpurun_out/$1_bench.log 2>&1; echo rc=$? >> gpurun_out/$1_bench.log; cat gpurun_out/$1_tests.log; python - <<PY
import json
for l in open("gpurun_out/$1_bench.log"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["mismatching_px"], d["parity"]["ok"])
    elif "rc=" in l: print(l)
PY
