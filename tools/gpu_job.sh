#!/bin/bash
# The GPU-box job of the next gpurun slot (edited between slots; every leg has its own timeout and log).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
leg() { # name timeout command...
  local name="$1" t="$2"; shift 2
  local t0=$(date +%s)
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "[$name] rc=$? $(( $(date +%s) - t0 ))s" | tee -a gpurun_out/job_summary.log
}
: > gpurun_out/job_summary.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee -a gpurun_out/job_summary.log
leg r02b_pt_cluster 400 python -m pytest tests/test_cluster_gpu.py tests/test_c2_scale_gpu.py -m gpu -q -s
leg r02b_pt_vae 500 python -m pytest tests/test_vae_gpu.py -m gpu -q
leg r02b_pt_tc 300 python -m pytest tests/test_tc_gpu.py tests/test_inputs.py -m gpu -q
leg r02b_pt_traj 400 python -m pytest tests/test_trajectory_gpu.py -m gpu -q -s
leg r02b_bench 760 bash -c 'python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02b_bench.json'
leg r02b_speed 200 env TC_MIN=128 python tools/train_speed.py
leg r02b_ncu 1500 bash tools/ncu_job.sh
for f in gpurun_out/r02b_pt_*.log; do echo "== $f"; grep -E "passed|failed|encode 1M|strict-RNG" $f | tail -4; done; cat gpurun_out/job_summary.log; grep "B=" gpurun_out/r02b_speed.log | cut -c1-60
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02b_bench.json"))
print({k: d[k] for k in ("value", "phases_s", "cluster_host_seconds", "clusters", "final_loss")})
print(d["e2e"]); print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline_cluster"]["achieved"], d["roofline_cluster"]["frac"])
PY
leg r02b_pt_cluster_ring6 400 env VK_PROBE_RING=6 python -m pytest tests/test_cluster_gpu.py -m gpu -q
leg r02b_pt_cluster_ring12 400 env VK_PROBE_RING=12 python -m pytest tests/test_cluster_gpu.py -m gpu -q -k 'golden or probe'
: > gpurun_out/r02b_probe_sweep.txt
for ring in 0 4 6 8 12; do VK_PROBE_RING=$ring timeout 120 python tools/probe_speed.py 2>&1 | grep "N=" | sed "s/^/RING=$ring /" >> gpurun_out/r02b_probe_sweep.txt; done
cat gpurun_out/r02b_probe_sweep.txt
