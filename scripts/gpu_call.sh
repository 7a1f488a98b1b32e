set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" 2>&1 | tail -4 > gpurun_out/t2.log
timeout 600 python scripts/bench_trainer.py --json gpurun_out/trainer.json > gpurun_out/trainer.log 2>&1
tail -3 gpurun_out/t2.log; cat gpurun_out/trainer.log
