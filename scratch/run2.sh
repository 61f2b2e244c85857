python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -3
for v in $VARIANTS; do IGD_FUSED_VARIANT=$v python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('variant $v', d['ms_per_step'], d['roofline']['frac'], d['roofline']['launch_ms'], d.get('parity_vs_oracle_on_timed_output'), d['clocks'])"; done
