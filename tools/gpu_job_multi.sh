#!/bin/bash
# Multi-GPU slot (gpurun --gpus N): the 2-GPU NCCL test, then the driver-style scaling run of bench.py at N GPUs.
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/job_multi_summary.log
leg() { local name="$1" t="$2"; shift 2; local t0=$(date +%s); timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
        echo "[$name] rc=$? $(( $(date +%s) - t0 ))s" | tee -a gpurun_out/job_multi_summary.log; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee -a gpurun_out/job_multi_summary.log
if [ "${TESTS:-1}" = "1" ]; then leg r02_pt_multigpu 400 python -m pytest tests/test_multigpu.py -m gpu -q; fi
leg r02_bench_dp${N} 800 bash -c "python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_dp${N}.json"
if [ "${REF:-1}" = "1" ]; then
leg r02_bench_ref_dp${N} 600 bash -c "python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_dp${N}.json"
fi
if [ "${WEAK:-0}" = "1" ]; then
leg r02_bench_dp${N}_weak 800 bash -c "python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --scaling weak --no-e2e > gpurun_out/r02_bench_dp${N}_weak.json"
fi
tail -3 gpurun_out/r02_pt_multigpu.log; tail -c 1200 gpurun_out/r02_bench_dp${N}.log; cat gpurun_out/job_multi_summary.log; head -c 600 gpurun_out/r02_bench_dp${N}.json; echo; head -c 400 gpurun_out/r02_bench_ref_dp${N}.json
