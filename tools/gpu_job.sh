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
leg r02p_pt_cluster 400 python -m pytest tests/test_cluster_gpu.py tests/test_c2_scale_gpu.py -m gpu -q
leg r02p_pt_vae 300 python -m pytest tests/test_vae_gpu.py tests/test_tc_gpu.py tests/test_trajectory_gpu.py -m gpu -q
leg r02p_probe_timeline 200 env VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/probe_timeline.py
leg r02p_clusterbench 300 python tools/cluster_speed.py
leg r02p_train_speed 300 env TC_MIN=128 python tools/train_speed.py
leg r02p_timeline 300 env VK_PDL=0 BATCHES=256 VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/kernel_timeline.py
leg r02p_bench 760 bash -c 'python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02p_bench.json'
for f in gpurun_out/r02p_pt_*.log; do echo "== $f"; grep -E "^FAILED|^ERROR|passed|failed" $f | tail -8; done
cat gpurun_out/job_summary.log; grep -E "N=|eval of|blocks" gpurun_out/r02p_probe_timeline.log; grep "lazy=" gpurun_out/r02p_clusterbench.log; grep -h "B=" gpurun_out/r02p_train_speed.log | cut -c1-250; grep -E "==|L[0-9]:|layer|backward" gpurun_out/r02p_timeline.log | cut -c1-250
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02p_bench.json"))
print({k: d[k] for k in ("value", "phases_s", "cluster_host_seconds", "clusters", "final_loss")})
print(d["e2e"]); print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline_cluster"]["achieved"], d["roofline_cluster"]["frac"])
PY
