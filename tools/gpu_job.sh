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
leg r02d_pt_cluster 400 python -m pytest tests/test_cluster_gpu.py tests/test_c2_scale_gpu.py -m gpu -q -s
leg r02d_pt_vae 500 python -m pytest tests/test_vae_gpu.py tests/test_tc_gpu.py tests/test_inputs.py -m gpu -q
leg r02d_pt_traj 400 python -m pytest tests/test_trajectory_gpu.py -m gpu -q -s
leg r02d_speed_r1 200 bash -c 'cd .r1 && TC_MIN=128 python tools/train_speed.py'
leg r02d_speed_now 200 env TC_MIN=128 python tools/train_speed.py
leg r02d_clusterbench 300 python tools/cluster_speed.py
leg r02d_graderr 400 env GRAD_CASES=50:4096,50:8192,100:512,100:2048 GRAD_OUT=gpurun_out/r02d_grad_error.json python tools/grad_error_fp64.py
leg r02d_bench 760 bash -c 'python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02d_bench.json'
for f in gpurun_out/r02d_pt_*.log; do echo "== $f"; grep -E "^FAILED|^ERROR|passed|failed|encode 1M|strict-RNG" $f | tail -8; done
cat gpurun_out/job_summary.log; grep "B=" gpurun_out/r02d_speed_*.log | cut -c1-90; tail -3 gpurun_out/r02d_clusterbench.log; grep -v WARNING gpurun_out/r02d_graderr.log | tail -30
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02d_bench.json"))
print({k: d[k] for k in ("value", "phases_s", "cluster_host_seconds", "clusters", "final_loss")})
print(d["e2e"]); print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline_cluster"]["achieved"], d["roofline_cluster"]["frac"])
PY
