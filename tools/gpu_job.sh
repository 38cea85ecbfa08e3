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
leg r02e_pt_cluster 400 python -m pytest tests/test_cluster_gpu.py -m gpu -q
leg r02e_pt_vae 500 python -m pytest tests/test_vae_gpu.py -m gpu -q
leg r02e_clusterbench 300 python tools/cluster_speed.py
leg r02e_timeline 300 env VK_PDL=0 VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/kernel_timeline.py
timeout 300 env CLUSTERS=4 ncu --set full --clock-control none --import-source on -k regex:eval_candidates -s 2 -c 2 -f -o gpurun_out/r02e_prof_eval python tools/prof_steps.py > gpurun_out/r02e_ncu_eval.log 2>&1
timeout 300 env CLUSTERS=4 VAMB_B200_CLUSTER_LAZY=0 ncu --set full --clock-control none --import-source on -k regex:eval_candidates -s 2 -c 2 -f -o gpurun_out/r02e_prof_eval_old python tools/prof_steps.py >> gpurun_out/r02e_ncu_eval.log 2>&1
leg r02e_bench 760 bash -c 'python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02e_bench.json'
for f in gpurun_out/r02e_pt_*.log; do echo "== $f"; grep -E "^FAILED|^ERROR|passed|failed" $f | tail -8; done
cat gpurun_out/job_summary.log; tail -3 gpurun_out/r02e_clusterbench.log; grep -E "==|L[0-9]:" gpurun_out/r02e_timeline.log | cut -c1-330
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02e_bench.json"))
print({k: d[k] for k in ("value", "phases_s", "cluster_host_seconds", "clusters", "final_loss")})
print(d["e2e"]); print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline_cluster"]["achieved"], d["roofline_cluster"]["frac"])
PY
