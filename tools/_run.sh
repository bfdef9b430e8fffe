python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/t19.log
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
for w in c5 c5v c5p c2 c3; do python bench.py --workload $w $Q > gpurun_out/b19_$w.json 2>gpurun_out/b19.err; tail -2 gpurun_out/b19.err; done
cat gpurun_out/t19.log
for f in gpurun_out/b19_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
