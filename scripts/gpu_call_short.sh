set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t2.log
timeout 300 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
timeout 600 python scripts/bench_rollout.py --json gpurun_out/rollout.json > gpurun_out/rollout.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_rollout.csv python scripts/bench_rollout.py --only TrainPhase2:16384:fused:24 --fixed > gpurun_out/ncu_rollout.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 2 -c 1 -o gpurun_out/prof_predict -f python scripts/predict_probe.py one 9 163840 0 > gpurun_out/ncu_predict.log 2>&1
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/t2.log; cut -c1-200 gpurun_out/bench_a.json; cat gpurun_out/rollout.log
