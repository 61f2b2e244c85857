mkdir -p gpurun_out
for v in $VARIANTS; do
IGD_FUSED_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o gpurun_out/prof_w$v python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_w$v.log 2>&1
tail -2 gpurun_out/ncu_w$v.log
done
