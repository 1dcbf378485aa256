#!/bin/sh
# Multi-host elastic run on a docker-compose "cluster": generate the compose file, build the CPU image, bring the cluster up and
# wait for the nodes (reference: benchmarks/adaptation/{build-image,compose-run}.sh).
#   benchmarks/adaptation/compose-run.sh [nodes=4] [slots=4]
# Without docker the same cast runs as plain processes on loopback aliases: python experiments/cluster_manager.py
set -e
cd "$(dirname "$0")/../.."
nodes=${1:-4}
slots=${2:-4}
docker/build-images.sh cpu
python benchmarks/adaptation/gen_compose.py --nodes "$nodes" --slots "$slots" > docker-compose.yaml
docker compose up --abort-on-container-exit
docker compose down
