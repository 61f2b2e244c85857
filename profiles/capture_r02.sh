#!/bin/bash
# Round-2 measurement recipe (B200_PROFILING.md): run on the GPU box, everything lands in gpurun_out/r02/.
#   gpurun --timeout 1500 -- 'bash profiles/capture_r02.sh'
# Then here (no GPU): python profiles/summarize.py r02 gpurun_out/r02   and   profiles/tools/linemix.py ...
set -u
O=gpurun_out/r02
mkdir -p $O
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
# 1) plain runs first (a number printed under ncu is never a bench value)
python bench.py > $O/bench.json 2> $O/bench.err || { echo "bench failed"; tail -5 $O/bench.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
python bench.py --cfg4 > $O/bench_cfg4_1gpu.json 2>> $O/bench.err
python bench.py --cfg5 > $O/bench_cfg5_1gpu.json 2>> $O/bench.err
python bench.py --chain --steps 10 > $O/chain_bench_shape.json 2>> $O/bench.err
python bench.py --chain --steps 10 --chain-bridges 16384 --chain-frames 100 > $O/chain_65536ch.json 2>> $O/bench.err
# 2) every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1
# 3) the top kernel, full set, with source
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 2 -o $O/prof_fused $CMD > $O/ncu_full.log 2>&1
# 4) the gateway chain: launch list + full capture of its fused kernel (packets in, packets out)
CH="python bench.py --chain --steps 2 --warmup 3 --chain-bridges 16384 --chain-frames 100"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 16 -c 16 --csv --log-file $O/chain_launches_65536ch.csv $CH > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 20 -c 15 --csv --log-file $O/chain_launches_bench_shape.csv python bench.py --chain --steps 2 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o $O/prof_gateway $CH > $O/ncu_gateway.log 2>&1
# 5) stand-alone kernels, sweeps, latencies
python profiles/kernel_rooflines.py > $O/kernel_rooflines.json 2> $O/kernel_rooflines.err
python profiles/tools/leg_count_sweep.py > $O/leg_count_sweep.txt 2>&1
python profiles/tools/gate_density_sweep.py > $O/gate_density_sweep.txt 2>&1
python profiles/tools/tick_latency.py > $O/tick_latency.txt 2>&1
python profiles/tools/pcie_probe.py > $O/pcie_probe.txt 2>&1
python profiles/tools/walks_bench.py > $O/walks_bench_shape.json 2>&1
python profiles/tools/walks_bench.py 16384 100 > $O/walks_65536ch.json 2>&1
ls -la $O | tail -30
