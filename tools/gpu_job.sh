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
leg r02c_pt_cluster 400 python -m pytest tests/test_cluster_gpu.py tests/test_c2_scale_gpu.py -m gpu -q -s
leg r02c_pt_vae 500 python -m pytest tests/test_vae_gpu.py -m gpu -q
leg r02c_pt_traj 400 python -m pytest tests/test_trajectory_gpu.py -m gpu -q -s
# same box: round-1 tree (git aeaa032, built in .r1/) against the current tree, step time per batch size
leg r02c_speed_r1 200 bash -c 'cd .r1 && TC_MIN=128 python tools/train_speed.py'
leg r02c_speed_now 200 env TC_MIN=128 python tools/train_speed.py
leg r02c_speed_now_tma0 200 env TC_MIN=128 VAMB_B200_TMA=0 python tools/train_speed.py
leg r02c_clusterbench 300 python tools/cluster_speed.py
leg r02c_pt_cluster_ring6 400 env VK_PROBE_RING=6 python -m pytest tests/test_cluster_gpu.py -m gpu -q -k "golden or probe or lazy"
: > gpurun_out/r02c_probe_sweep.txt
for ring in 0 4 6 8 12; do VK_PROBE_RING=$ring timeout 120 python tools/probe_speed.py 2>&1 | grep "N=" | sed "s/^/RING=$ring /" >> gpurun_out/r02c_probe_sweep.txt; done
leg r02c_ncu 1500 bash tools/ncu_job.sh
for f in gpurun_out/r02c_pt_*.log; do echo "== $f"; grep -E "^FAILED|^ERROR|passed|failed|encode 1M|strict-RNG" $f | tail -8; done
cat gpurun_out/job_summary.log; cat gpurun_out/r02c_probe_sweep.txt; grep "B=" gpurun_out/r02c_speed_*.log | cut -c1-90; cat gpurun_out/r02c_clusterbench.log | tail -12
