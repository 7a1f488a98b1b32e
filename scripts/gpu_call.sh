set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t2.log
timeout 300 python scripts/bench_rollout.py --json gpurun_out/rollout.json > gpurun_out/rollout.log 2>&1
timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
CA_STORE_MODE=vec4 timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_vec4.json 2>> gpurun_out/bench_a.err
CA_STEP_KERNEL=pipe timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_pipe.json 2>> gpurun_out/bench_a.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 2 -c 1 -o gpurun_out/prof_predict_r01 -f python scripts/predict_probe.py one 9 163840 0 > gpurun_out/ncu_predict.log 2>&1
tail -5 gpurun_out/t2.log; cat gpurun_out/rollout.log; cut -c1-200 gpurun_out/bench_a.json gpurun_out/bench_vec4.json gpurun_out/bench_pipe.json
