"""Fill the rollout table of DESIGN.md from the JSON written by scripts/bench_rollout.py --json."""
import json
import sys

rows = json.load(open(sys.argv[1]))
key = {("TrainPhase2", True, "fused"): "P2_FIXED", ("TrainPhase2", False, "fused"): "P2_MIX", ("TrainPhase2", False, "composed"): "P2_COMP",
       ("TrainPhase1", True, "fused"): "P1_FIXED", ("TrainPhase1", False, "fused"): "P1_MIX", ("TrainPhase1", False, "composed"): "P1_COMP"}
s = open("DESIGN.md").read()
for r in rows:
    k = key[(r["config"], bool(r["fixed_agents"]), r["predictor"])]
    s = s.replace("ROLL_%s_MS" % k, "%.3f" % r["rollout_ms_per_step"])
    s = s.replace("ROLL_%s_AS" % k, "%.0f (%.0f live, %.0f learning per step)" % (r["agent_steps_per_s"] / 1e6, r["live_agents_per_step"], r["learning_agents_per_step"]))
    s = s.replace("ROLL_%s_ROWS" % k, "%.0f" % (r["learner_rows_per_s"] / 1e6))
open("DESIGN.md", "w").write(s)
