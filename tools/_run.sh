python -m pytest tests -x -q -m gpu 2>&1 | tail -3
Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py --steps 200 --workload m64 $Q > gpurun_out/b77.json 2>>gpurun_out/b77.err; python -c "
import json
d=json.load(open('gpurun_out/b77.json'))
print('m64 value %.3e ms/step %.4f frac %.3f | flushed ms %.4f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median']))"
tail -2 gpurun_out/b77.err
