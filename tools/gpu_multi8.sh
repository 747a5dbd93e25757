#!/bin/bash
TAG=${1:-multi8}; N=${2:-8}; O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench$N exit $?"; tail -c 300 $O/bench_n$N.err
python - <<PY
import json
d=json.load(open("$O/bench_n$N.json"))
print("value",d["value"],"e2e",d["e2e"]["value"], d["e2e"]["ms_per_step"])
g=d["ba"]["global_ba"]; print("global_ba", g["value"], g["ms_per_solve"], g.get("vs_1gpu"))
PY
