#!/bin/bash
# Final check of a round, as the driver runs it: the whole `-m gpu` suite, smoke(), then the default bench command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/job_final_summary.log
leg() { local name="$1" t="$2"; shift 2; local t0=$(date +%s); timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
        echo "[$name] rc=$? $(( $(date +%s) - t0 ))s" | tee -a gpurun_out/job_final_summary.log; }
leg r02_final_pytest 900 python -m pytest tests/ -x -q -m gpu
leg r02_final_smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
leg r02_final_bench 800 bash -c 'python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json'
tail -3 gpurun_out/r02_final_pytest.log; tail -2 gpurun_out/r02_final_smoke.log; cat gpurun_out/job_final_summary.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_final_bench.json"))
print({k: d[k] for k in ("value", "phases_s", "clusters", "final_loss", "gpu_launches", "clocks")})
print(d["e2e"]); print(d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["traffic"]); print(d["roofline_cluster"]["achieved"], d["roofline_cluster"]["frac"]); print(d["cpu_baseline"]["value"])
PY
