mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_predictor.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t2.log
timeout 300 python scripts/predict_probe.py > gpurun_out/predict_probe.log 2>&1
tail -3 gpurun_out/t2.log; grep -A1 "^M=" gpurun_out/predict_probe.log
