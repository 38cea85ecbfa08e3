#!/bin/bash
# ncu leg of a GPU slot (one GPU; profiler numbers are never bench values): launch list + full captures of the
# dominant kernels.  Reports land in gpurun_out/ (kept small: gpurun merges at most 64 MiB back), summaries are
# extracted on the CPU box by tools/ncu_extract.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python tools/prof_steps.py > gpurun_out/r02_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'probe_kernel|eval_candidates' -s 4 -c 4 -f -o gpurun_out/r02_prof_cluster \
    python tools/prof_steps.py > gpurun_out/r02_ncu_cluster.log 2>&1
# training kernels (a --set full report is ~2.4 MB per launch and gpurun merges at most 64 MiB back): at B = 4096 the six
# forward launches and the first three backward ones (layers 5, 4, 3) of the first step ...
timeout 400 env STEPS=1 CLUSTERS=0 ncu --set full --clock-control none \
    -k regex:'layer_tc_kernel' -c 9 -f -o gpurun_out/r02_prof_vae4096 \
    python tools/prof_steps.py > gpurun_out/r02_ncu_vae.log 2>&1
# ... and the same nine at B = 256 (the third step of the run), with source correlation
timeout 400 env STEPS=1 CLUSTERS=0 ncu --set full --clock-control none --import-source on \
    -k regex:'layer_tc_kernel' -s 24 -c 9 -f -o gpurun_out/r02_prof_vae256 \
    python tools/prof_steps.py >> gpurun_out/r02_ncu_vae.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
