set -x
mkdir -p gpurun_out
./build/mufu_probe > gpurun_out/mufu_probe.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t2.log
timeout 600 python scripts/bench_rollout.py --json gpurun_out/rollout.json > gpurun_out/rollout.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_rollout.csv python scripts/bench_rollout.py --only TrainPhase2:16384:fused:24 --fixed > gpurun_out/ncu_rollout.log 2>&1
cat gpurun_out/mufu_probe.log; tail -5 gpurun_out/t2.log; cat gpurun_out/rollout.log
