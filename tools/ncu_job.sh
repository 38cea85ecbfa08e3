#!/bin/bash
# ncu leg of a GPU slot (one GPU; profiler numbers are never bench values): launch list + full captures of the
# dominant kernels.  Reports land in gpurun_out/, summaries are extracted on the CPU box by tools/ncu_extract.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export VAMB_B200_TMA=${VAMB_B200_TMA:-0} VAMB_B200_WGRAD_FLUSH=${VAMB_B200_WGRAD_FLUSH:-0}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python tools/prof_steps.py > gpurun_out/r02_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:probe_kernel -c 3 -f -o gpurun_out/r02_prof_probe \
    python tools/prof_steps.py > gpurun_out/r02_ncu_probe.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:eval_candidates -c 3 -f -o gpurun_out/r02_prof_eval \
    python tools/prof_steps.py > gpurun_out/r02_ncu_eval.log 2>&1
# training kernels: the first step at B = 4096 (12 layer launches + loss + optimiser) and one at B = 256
timeout 600 env STEPS=1 CLUSTERS=0 ncu --set full --clock-control none --import-source on \
    -k regex:'layer_tc_kernel|loss_kernel|dadapt_kernel' -c 42 -f -o gpurun_out/r02_prof_vae \
    python tools/prof_steps.py > gpurun_out/r02_ncu_vae.log 2>&1
ls -la gpurun_out/*.ncu-rep
