#!/bin/bash
# Same call surface as the reference's train.sh:9-15:  ./train.sh [ConfigClass] [config path]
set -e
DIR="$( cd "$( dirname "${BASH_SOURCE[0]}" )" && pwd )"
export GYM_CONFIG_CLASS=${1:-TrainPhase2}
export GYM_CONFIG_PATH=${2:-$DIR/rl_collision_avoidance_b200/ga3c/Config.py}
cd "$DIR"
NGPU=${NGPU:-1}
if [ "$NGPU" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NGPU" --master-addr 127.0.0.1 --master-port ${MASTER_PORT:-29511} \
    -m rl_collision_avoidance_b200.ga3c.Run
else
  python -m rl_collision_avoidance_b200.ga3c.Run
fi
