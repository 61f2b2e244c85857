#!/bin/bash
# Round 2, second capture (lane walks): run on the GPU box, everything lands in gpurun_out/r02b/.
#   gpurun --timeout 1200 -- 'bash profiles/capture_r02b.sh'
set -u
O=gpurun_out/r02b
mkdir -p $O
# 1) plain runs first (a number printed under ncu is never a bench value)
python bench.py > $O/bench.json 2> $O/bench.err || { echo "bench failed"; tail -5 $O/bench.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
python bench.py --chain --steps 20 > $O/chain_bench_shape.json 2>> $O/bench.err
python bench.py --chain --steps 20 --serial-walks --no-e2e > $O/chain_bench_shape_serial_walks.json 2>> $O/bench.err
python bench.py --chain --steps 20 --chain-bridges 16384 --chain-frames 100 > $O/chain_65536ch.json 2>> $O/bench.err
python bench.py --chain --steps 20 --no-e2e --chain-bridges 4096 --chain-frames 400 > $O/chain_16384ch_400ticks.json 2>> $O/bench.err
python bench.py --chain --steps 20 --no-e2e --serial-walks --chain-bridges 4096 --chain-frames 400 > $O/chain_16384ch_400ticks_serial_walks.json 2>> $O/bench.err
python bench.py --chain --steps 20 --no-e2e --chain-bridges 8192 --chain-frames 200 > $O/chain_32768ch_200ticks.json 2>> $O/bench.err
python bench.py --chain --steps 20 --no-e2e --serial-walks --chain-bridges 8192 --chain-frames 200 > $O/chain_32768ch_200ticks_serial_walks.json 2>> $O/bench.err
# 2) the gateway call: every launch with its device time (cold-cache, serialised: compare SHARES), lane walks and serial walks
CH="python bench.py --chain --no-e2e --steps 2 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 80 --csv --log-file $O/chain_launches_bench_shape.csv $CH > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $O/chain_launches_bench_shape_serial_walks.csv $CH --serial-walks > /dev/null 2>&1
# 3) the walk kernels, full set, with source
ncu --set full --clock-control none --import-source on -k regex:k_rxarb -s 3 -c 1 -o $O/prof_rxarb $CH > $O/ncu_rxarb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_plan_walk -s 3 -c 1 -o $O/prof_plan $CH > $O/ncu_plan.log 2>&1
ls -la $O | tail -30
