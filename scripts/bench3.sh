#!/bin/bash
# development aid: the three execution modes of bench.py back to back (graph / Python launches / host lengths)
tag=${1:-x}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -5 gpurun_out/${tag}_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${tag}_bench_nograph.json 2>> gpurun_out/${tag}_bench.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --host-lengths > gpurun_out/${tag}_bench_host.json 2>> gpurun_out/${tag}_bench.err
python - <<EOF2
import json
for f in ("${tag}_bench","${tag}_bench_nograph","${tag}_bench_host"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d["gpu_launches"], d["parity"]["k_gpu"], d["parity"]["logit_max_abs"], d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", e)
EOF2
