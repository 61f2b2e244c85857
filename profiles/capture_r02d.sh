#!/bin/bash
# Round 2, last runs (shipped binary): what produced profiles/bench_r02d*.json, bench_r02c*.json, r02d_rxarb_*.txt,
# r02c_rxarb_bridge_ncu.txt and r02c_chain_launches_65536ch.txt.   gpurun --timeout 1200 -- 'bash profiles/capture_r02d.sh'
set -u
O=gpurun_out/r02d
mkdir -p $O
# plain runs first (a number printed under ncu is never a bench value)
python bench.py > $O/bench.json 2> $O/bench.err || { echo "bench failed"; tail -5 $O/bench.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
python bench.py --chain --steps 20 > $O/chain_bench_shape.json 2>> $O/bench.err
python bench.py --chain --steps 20 --no-e2e --chain-bridges 4096 --chain-frames 400 > $O/chain_16384ch_400ticks.json 2>> $O/bench.err
python bench.py --chain --steps 20 --no-e2e --chain-bridges 8192 --chain-frames 200 > $O/chain_32768ch_200ticks.json 2>> $O/bench.err
python bench.py --chain --steps 20 --chain-bridges 16384 --chain-frames 100 > $O/chain_65536ch.json 2>> $O/bench.err
python bench.py --chain --steps 20 --no-e2e --chain-bridges 16384 --chain-frames 200 > $O/chain_65536ch_200ticks.json 2>> $O/bench.err
python profiles/tools/gateway_tick_latency.py > $O/gateway_tick_latency.json 2>> $O/bench.err
# launch list of one wide gateway call, then the two receive-side walk kernels with the full set and source
W="python bench.py --chain --no-e2e --steps 2 --warmup 3 --chain-bridges 16384 --chain-frames 100"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file $O/chain_launches_65536ch.csv $W > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rxarb_bridge -s 3 -c 1 -o $O/prof_rxarb_bridge $W > $O/ncu_bridge.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rxarb_walk -s 3 -c 1 -o $O/prof_rxarb python bench.py --chain --no-e2e --steps 2 --warmup 3 > $O/ncu_rxarb.log 2>&1
# here (no GPU): python profiles/tools/ncusum.py $O/prof_rxarb.ncu-rep "" 53248 ; python profiles/tools/linemix.py $O/prof_rxarb.ncu-rep igate4xsoftphonedsp_b200/libigate_dsp.so k_rxarb_walk 53248
ls -la $O | tail -20
