set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/t2.log
timeout 300 python scripts/bench_rollout.py --json gpurun_out/rollout.json > gpurun_out/rollout.log 2>&1
timeout 300 python bench.py --steps 1200 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
CA_STEP_KERNEL=pipe timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_pipe.json 2>> gpurun_out/bench_a.err
timeout 300 python scripts/predict_probe.py > gpurun_out/predict_probe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 2 -c 1 -o gpurun_out/prof_predict_r01 -f python scripts/predict_probe.py one 9 163840 0 > gpurun_out/ncu_predict.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ca_step_kernel -s 30 -c 2 -o gpurun_out/prof_step_r01 -f python bench.py --steps 48 --warmup 12 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 96 --warmup 12 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -5 gpurun_out/t2.log; cat gpurun_out/rollout.log; cat gpurun_out/predict_probe.log; cut -c1-300 gpurun_out/bench_a.json gpurun_out/bench_pipe.json
