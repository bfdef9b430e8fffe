python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/t8.log
python bench.py --steps 200 --warmup 5 > gpurun_out/b8_c5.json 2>gpurun_out/b8.err
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py --workload c3 $Q > gpurun_out/b8_c3.json 2>>gpurun_out/b8.err
cat gpurun_out/t8.log; tail -3 gpurun_out/b8.err
for f in gpurun_out/b8_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s | e2e %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value'], d['e2e']))
"; done
