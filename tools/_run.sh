python -m pytest tests -m gpu -q 2>&1 | tail -6
Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py --steps 200 $Q > gpurun_out/b35_c5.json 2>gpurun_out/b35.err
python bench.py --steps 50 $Q > gpurun_out/b35_c5_50.json 2>>gpurun_out/b35.err
python bench.py --steps 20 $Q > gpurun_out/b35_c5_20.json 2>>gpurun_out/b35.err
python bench.py --steps 3 $Q > gpurun_out/b35_c5_3.json 2>>gpurun_out/b35.err
python bench.py --steps 200 --workload c4 $Q > gpurun_out/b35_c4.json 2>>gpurun_out/b35.err
tail -2 gpurun_out/b35.err
for f in gpurun_out/b35_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s | launches %d %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value'], d['gpu_launches'], d['config']['launch']))
"; done
