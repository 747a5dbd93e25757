#!/bin/bash
TAG=${1:-multi2}; O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ba_multi_check.py > $O/multi_check.log 2>&1; echo "multi exit $?"; grep -E "keyframes:|MULTI|rejected" $O/multi_check.log | grep -E "rank 0|MULTI|rejected"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 exit $?"; tail -c 300 $O/bench_n2.err
python - <<PY
import json
d=json.load(open("$O/bench_n2.json"))
print("value",d["value"],"e2e",d["e2e"]["value"])
g=d["ba"]["global_ba"]; print("global_ba", g["value"], g["ms_per_solve"], g.get("vs_1gpu"))
PY
