python -m pytest tests -m gpu -q 2>&1 | tail -4
Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py --steps 200 $Q > gpurun_out/b34_c5.json 2>gpurun_out/b34.err
python bench.py --steps 1000 $Q > gpurun_out/b34_c5_1000.json 2>>gpurun_out/b34.err
python bench.py --steps 50 $Q > gpurun_out/b34_c5_50.json 2>>gpurun_out/b34.err
python bench.py --steps 200 --workload c3 $Q > gpurun_out/b34_c3.json 2>>gpurun_out/b34.err
tail -2 gpurun_out/b34.err
for f in gpurun_out/b34_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
