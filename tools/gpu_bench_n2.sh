#!/bin/bash
# two-rank bench line (one process per GPU, NCCL barrier + max-reduction of the elapsed time only)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 300 gpurun_out/r2_bench_n2.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n2.json") if l.startswith("{")][-1])
print(d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], d["config"]["host_wait"], d["config"]["host_cores"], d["clocks"])
PY
