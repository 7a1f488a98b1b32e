set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_predictor.py tests/test_gpu_ga3c.py tests/test_gpu_scenarios.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t2.log
timeout 300 python scripts/predict_probe.py > gpurun_out/predict_probe.log 2>&1
timeout 300 python scripts/bench_rollout.py --only TrainPhase2:16384:fused:60 --fixed > gpurun_out/rollout.log 2>&1
tail -5 gpurun_out/t2.log; cat gpurun_out/predict_probe.log gpurun_out/rollout.log
