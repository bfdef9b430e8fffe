python -m pytest tests -m gpu -q 2>&1 | tail -4
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
for w in c5 c4 c3 c2 c5v; do python bench.py --workload $w $Q > gpurun_out/b33_$w.json 2>gpurun_out/b33.err; tail -1 gpurun_out/b33.err; done
for f in gpurun_out/b33_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
