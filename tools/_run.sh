python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/t25.log
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py --workload c4 $Q > gpurun_out/b25_c4.json 2>gpurun_out/b25.err
python bench.py --workload c4 --envs 8 $Q > gpurun_out/b25_c4_e8.json 2>>gpurun_out/b25.err
cat gpurun_out/t25.log; tail -3 gpurun_out/b25.err
for f in gpurun_out/b25_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | launches %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['gpu_launches']))
"; done
