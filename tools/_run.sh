python -m pytest tests -m gpu -q 2>&1 | tail -4
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py --workload c4 $Q > gpurun_out/b32_c4.json 2>gpurun_out/b32.err
tail -2 gpurun_out/b32.err
python -c "
import json,sys
d=json.load(open('gpurun_out/b32_c4.json'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac']))
"
BA="--workload c4 --steps 20 --warmup 3 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $BA > gpurun_out/plain32.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4_new3.csv python bench.py $BA > gpurun_out/ncu32.log 2>&1
