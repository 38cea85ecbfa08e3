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
leg r02g_pt_vae 500 python -m pytest tests/test_vae_gpu.py tests/test_trajectory_gpu.py -m gpu -q
leg r02g_probe_timeline 200 env VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/probe_timeline.py
leg r02g_timeline 300 env VK_PDL=0 VAMB_B200_SO=vamb_b200/_vk_timeline.so python tools/kernel_timeline.py
leg r02g_speed_now 200 env TC_MIN=128 python tools/train_speed.py
leg r02g_speed_r1 200 bash -c 'cd .r1 && TC_MIN=128 python tools/train_speed.py'
for f in gpurun_out/r02g_pt_*.log; do echo "== $f"; grep -E "^FAILED|^ERROR|passed|failed" $f | tail -8; done
cat gpurun_out/job_summary.log; grep -E "N=|eval of" gpurun_out/r02g_probe_timeline.log; grep -E "==|L[0-9]:" gpurun_out/r02g_timeline.log | cut -c1-300; grep "B=" gpurun_out/r02g_speed_*.log | cut -c1-100
