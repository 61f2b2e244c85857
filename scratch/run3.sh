for rep in 1 2; do
for lib in scratch/lib_xs4.so scratch/lib_xs4p1.so scratch/lib_xs3p1.so; do IGD_LIB_PATH=${lib:+$PWD/$lib} python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('lib=$lib', d['ms_per_step'], d['roofline']['frac'], d['roofline']['launch_ms']['min'], d.get('parity_vs_oracle_on_timed_output'), d['clocks']['sm_mhz'])"; done; done
