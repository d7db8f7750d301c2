#!/bin/bash
# tools/final_run3_light.sh TAG: bench.py (full line) + the ncu launch list of a short run of the same workload
# (plain launches: NF_MG_GRAPH=0), after that short run has exited 0 without ncu.  The heavy passes are in final_run3.sh.
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_4097.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
NF_MG_GRAPH=0 $SHORT > $O/${TAG}_short.json 2>> $O/${TAG}_bench.err; echo "short rc=$?"
NF_MG_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv \
    $SHORT > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launches.txt 2>&1; head -16 $O/${TAG}_launches.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
