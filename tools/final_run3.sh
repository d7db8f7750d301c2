#!/bin/bash
# tools/final_run3.sh TAG: the single-GPU evidence of a round, in the order the profiling recipe asks for (every ncu pass
# only after the same command has exited 0 without ncu):
#   1. bench.py (full line: roofline, cpu_baseline with the port's 4097^2 kernel times, e2e)            -> ${TAG}_bench_4097.json
#   2. bench.py --impl reference                                                                        -> ${TAG}_bench_reference_arm.json
#   3. bench.py --grid 16385 --steps 20 (BASELINE.json config 5 on one GPU)                             -> ${TAG}_bench_n1_16385.json
#   4. ncu launch list of a short bench run (plain launches: NF_MG_GRAPH=0, same kernels)               -> ${TAG}_launches.txt
#   5. ncu --set full of the finest-level smoother launches of that run                                 -> ${TAG}_prof_stream.ncu-rep
#   6. ncu metrics pass over every kernel of the library (tools/ncu_kernels_run.py)                     -> ${TAG}_ncu_all_kernels_4097.txt
TAG=${1:-r2j}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_4097.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2>> $O/${TAG}_bench.err; echo "reference arm rc=$?"
python bench.py --grid 16385 --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_n1_16385.json 2>> $O/${TAG}_bench.err; echo "16385 rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
NF_MG_GRAPH=0 $SHORT > $O/${TAG}_short.json 2>> $O/${TAG}_bench.err; echo "short rc=$?"
NF_MG_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv \
    $SHORT > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launches.txt 2>&1; head -12 $O/${TAG}_launches.txt
NF_MG_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:k_rbsor_stream -s 40 -c 8 -f \
    -o $O/${TAG}_prof_stream $SHORT > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/${TAG}_prof_stream.ncu-rep --page details > $O/${TAG}_ncu_k_rbsor_stream_details_raw.txt 2>&1
ncu -i $O/${TAG}_prof_stream.ncu-rep --page raw --csv > $O/${TAG}_ncu_k_rbsor_stream_raw.csv 2>&1
python tools/ncu_kernels_run.py 4097 > $O/${TAG}_kernels_plain.log 2>&1; echo "kernels plain rc=$?"
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $METRICS --clock-control none --csv --log-file $O/${TAG}_ncu_all.csv python tools/ncu_kernels_run.py 4097 \
    > $O/${TAG}_ncu_all.log 2>&1; echo "ncu all rc=$?"
python tools/summarize_ncu.py $O/${TAG}_ncu_all.csv > $O/${TAG}_ncu_all_kernels_4097.txt 2>&1; head -30 $O/${TAG}_ncu_all_kernels_4097.txt
