#!/bin/bash
# tools/final_run.sh TAG: the single-GPU evidence run of a round -- GPU tests, smoke, bench, config-4 sweep, the ncu launch list
# of the bench command and one `--set full` capture of the dominant kernel (all written under gpurun_out/).
TAG=${1:-r1f}
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.log
python bench.py > $O/${TAG}_bench_4097.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2>> $O/${TAG}_bench.err; echo "ref rc=$?"
python tools/run_configs.py c4 2049 > $O/${TAG}_config4_sweep_2049.json 2> $O/${TAG}_c4.err; echo "c4 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launches.txt 2>&1; head -12 $O/${TAG}_launches.txt
ncu --set full --clock-control none --import-source on -k regex:k_rbsor_tma -s 30 -c 3 -f -o $O/prof_rbsor_tma_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
