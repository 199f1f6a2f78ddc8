#!/bin/bash
# Round-end evidence on ONE GPU box: GPU tests, smoke, the five bench configurations, the CPU arm, the ncu launch list
# of a bench run and the `ncu --set full` capture of scripts/ncu_forward.py. Everything lands in gpurun_out/<tag>_*.
tag=${1:-r2j}
o=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $o/${tag}_gpu_tests.log 2>&1; tail -3 $o/${tag}_gpu_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -2 $o/${tag}_bench.err
for c in 1 3 4 5; do
  python bench.py --config $c --steps 20 --warmup 5 > $o/${tag}_bench_c$c.json 2> $o/${tag}_bench_c$c.err; tail -2 $o/${tag}_bench_c$c.err
done
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
python bench.py --steps 20 --warmup 5 --streams 1 --no-cpu-baseline > $o/${tag}_bench_s1.json 2> $o/${tag}_bench_s1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 3 --streams 1 --no-cpu-baseline > $o/${tag}_ncu_bench.log 2>&1; echo "ncu launches exit=$?"; wc -l $o/${tag}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/${tag}_full \
  python scripts/ncu_forward.py > $o/${tag}_ncu_full.log 2>&1; echo "ncu full exit=$?"
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > $o/${tag}_ncu_full_raw.csv 2>/dev/null; wc -l $o/${tag}_ncu_full_raw.csv
python - <<EOF2
import json
for f in ("bench", "bench_c1", "bench_c3", "bench_c4", "bench_c5", "bench_s1", "bench_reference"):
    try:
        d = json.load(open("$o/${tag}_%s.json" % f))
        r = d.get("roofline") or {}
        print(f, round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), r.get("kernel"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
EOF2
