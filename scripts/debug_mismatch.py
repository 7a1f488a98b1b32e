"""Debug helper (GPU box): run CUDA vs oracle on seeded random worlds and print every mismatch with context."""
import sys
import numpy as np
sys.path.insert(0, ".")
from rl_collision_avoidance_b200 import _abi
from rl_collision_avoidance_b200.vec_env import HostVecEnv
from oracle.ca_oracle import OracleEnv
from tests.test_gpu_parity import _random_worlds

A, M, W, side, sort, seed_off = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), sys.argv[5], 0
rng = np.random.default_rng(1000 + 7 * A + M)
init, nag = _random_worlds(rng, W, A, side, policies=(0, 0, 0, 1, 2), ragged=True)
cfg = _abi.default_config(W, A, M, sort_method=_abi.SORT_METHODS[sort])
gpu, cpu = HostVecEnv(cfg, want_sorted_idx=True), OracleEnv(cfg)
gpu.set_world_state(init, nag); cpu.set_world_state(init, nag)
gpu.reset(); cpu.reset()
np.set_printoptions(precision=17, linewidth=200)
for t in range(70):
    act = rng.choice([0, 1, 2, 2, 2, 3, 4, 5, 6, 7, 8, 9, 10], size=(W, A)).astype(np.int32)
    gs0, cs0 = gpu.get_state(), cpu.get_state()
    gpu.step(act); cpu.step(act)
    bad = np.argwhere(np.abs(gpu.obs - cpu.obs) > 1e-5)
    badflag = np.argwhere(gpu.done != cpu.done)
    badidx = np.argwhere(gpu.sorted_idx != cpu.sorted_idx)
    if len(bad) or len(badflag) or len(badidx):
        print("t=%d: %d obs mismatches, %d done mismatches, %d idx mismatches" % (t, len(bad), len(badflag), len(badidx)))
        for w, i, c in bad[:6]:
            print(" obs w=%d i=%d col=%d n=%d gpu=%r cpu=%r policy=%s" % (w, i, c, nag[w], gpu.obs[w, i, c], cpu.obs[w, i, c], init[w, :, _abi.I_POLICY]))
            print("   pre-state gpu:", gs0[w, i]); print("   pre-state cpu:", cs0[w, i])
            print("   post-state gpu:", gpu.get_state()[w, i]); print("   post-state cpu:", cpu.get_state()[w, i])
            print("   action", act[w, i])
        for w, i in badflag[:6]:
            print(" done w=%d i=%d gpu=%d cpu=%d" % (w, i, gpu.done[w, i], cpu.done[w, i]))
        for w, i, k in badidx[:6]:
            print(" idx w=%d i=%d slot=%d gpu=%s cpu=%s" % (w, i, k, gpu.sorted_idx[w, i], cpu.sorted_idx[w, i]))
        break
else:
    print("no mismatch in 70 steps")
