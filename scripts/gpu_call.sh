set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ga3c.py tests/test_gpu_scenarios.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t2.log
timeout 600 python scripts/bench_rollout.py --json gpurun_out/rollout.json > gpurun_out/rollout.log 2>&1
tail -4 gpurun_out/t2.log; cat gpurun_out/rollout.log
