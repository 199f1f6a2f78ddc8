#!/bin/bash
# Multi-GPU bench lines on one 8-GPU box: configuration 2 at N = 8, configuration 5 (BLIP-VQA, BASELINE: 8 GPUs) at
# N = 8, configuration 4 (CLIP, BASELINE: 4 GPUs) at N = 4.  usage: scripts/multi_gpu.sh <tag>
tag=${1:-r2j}
o=gpurun_out
run() {  # config, N
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29600 + $1 * 10 + $2)) \
    bench.py --config $1 --gpus $2 --steps 20 --warmup 5 > $o/${tag}_bench_c$1_n$2.json 2> $o/${tag}_bench_c$1_n$2.err
  python - <<EOF2
import json
try:
    d = json.load(open("$o/${tag}_bench_c$1_n$2.json"))
    print("config $1 N=$2", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), [round(v, 2) for v in d["ms_per_step_by_rank"]], d["clocks"])
except Exception as e:
    print("config $1 N=$2 ERR", e)
EOF2
}
python bench.py --no-cpu-baseline --steps 20 --warmup 5 > $o/${tag}_bench_c2_n1.json 2>/dev/null
python -c "
import json; d=json.load(open('$o/${tag}_bench_c2_n1.json')); print('config 2 N=1', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks'])"
run 2 8
run 5 8
run 4 4
