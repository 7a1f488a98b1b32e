"""TEST INFRASTRUCTURE — stages the Python sources of the reference's env path under the git-ignored oracle/_ref/.

    python oracle/stage_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

/root/reference does not exist on the GPU box, but oracle/_ref/ travels there with the gpurun snapshot (it is
git-ignored, not gpurun-ignored), so bench.py can time the UNMODIFIED reference env on the box's host cores
(oracle/ref_bench.py, `cpu_baseline_reference` in the bench line).  Only *.py files of
gym-collision-avoidance/gym_collision_avoidance are copied (no checkpoints, maps or pickles: the timed workload is
get_testcase_random, which needs none); nothing under oracle/_ref/ is ever committed or imported by the product."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CA_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
PKG = os.path.join("gym-collision-avoidance", "gym_collision_avoidance")


def stage(verbose=True):
    src_pkg = os.path.join(SRC, PKG)
    if not os.path.isdir(src_pkg):
        if verbose:
            print("stage_ref: %s not present, nothing staged" % src_pkg)
        return 0
    n = 0
    for root, dirs, files in os.walk(src_pkg):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "checkpoints", "world_maps", "test_cases", "logs")]
        rel = os.path.relpath(root, SRC)
        for f in files:
            if f.endswith(".py"):
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel, f))
                n += 1
    if verbose:
        print("stage_ref: %d reference .py files under %s" % (n, DST))
    return n


if __name__ == "__main__":
    sys.exit(0 if stage() >= 0 else 1)
