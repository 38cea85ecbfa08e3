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
# known-good paths first, experimental variants (TMA operand, wgrad flush) in their own processes: a trapped kernel
# poisons only its own CUDA context
leg r02_pt_cluster 400 python -m pytest tests/test_cluster_gpu.py tests/test_c2_scale_gpu.py -m gpu -q
leg r02_pt_vae 500 python -m pytest tests/test_vae_gpu.py -m gpu -q -k "not tma and not flush"
leg r02_pt_tc 300 python -m pytest tests/test_tc_gpu.py tests/test_inputs.py -m gpu -q -k "not tma and not flush"
leg r02_bench 760 bash -c 'python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench.json'
leg r02_ref 300 bash -c 'python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_ref.json'
leg r02_pt_traj 400 python -m pytest tests/test_trajectory_gpu.py -m gpu -q
leg r02_pt_tc_tma 300 python -m pytest tests/test_tc_gpu.py -m gpu -q -k "tma"
leg r02_pt_tc_flush 300 python -m pytest tests/test_tc_gpu.py -m gpu -q -k "flush"
leg r02_pt_vae_tma 400 python -m pytest tests/test_vae_gpu.py -m gpu -q -k "tma"
leg r02_pt_vae_flush 400 python -m pytest tests/test_vae_gpu.py -m gpu -q -k "flush"
leg r02_tf32 120 python tools/measure_tf32_peak.py
leg r02_graderr 300 python tools/grad_error_fp64.py
leg r02_speed_tma0 200 env TC_MIN=128 VAMB_B200_TMA=0 python tools/train_speed.py
leg r02_speed_tma1 200 env TC_MIN=128 VAMB_B200_TMA=1 python tools/train_speed.py
for f in gpurun_out/r02_pt_*.log; do echo "== $f"; tail -4 $f; done; cat gpurun_out/job_summary.log; tail -c 1500 gpurun_out/r02_bench.log
: > gpurun_out/r02_probe_sweep.txt
for cfg in "4 4" "4 2" "4 6" "4 8" "8 3" "8 2" "8 4"; do set -- $cfg; VK_PROBE_R=$1 VK_PROBE_BPS=$2 timeout 120 python tools/probe_speed.py >> gpurun_out/r02_probe_sweep.txt 2>&1; done
cat gpurun_out/r02_probe_sweep.txt | grep "N=" 
