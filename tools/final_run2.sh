#!/bin/bash
# tools/final_run2.sh TAG: bench (full line), 16385^2 on one GPU, ncu launch list + full capture of the dominant kernel
TAG=${1:-r1g}
O=gpurun_out
python bench.py > $O/${TAG}_bench_4097.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --grid 16385 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/${TAG}_bench_n1_16385.json 2>> $O/${TAG}_bench.err; echo "16385 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launches.txt 2>&1; head -8 $O/${TAG}_launches.txt
ncu --set full --clock-control none --import-source on -k regex:k_rbsor_tma -s 30 -c 2 -f -o $O/prof_rbsor_tma_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
