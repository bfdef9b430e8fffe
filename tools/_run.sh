python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0 > gpurun_out/b52_c5.json 2>gpurun_out/b52.err; python -c "
import json
d=json.load(open('gpurun_out/b52_c5.json'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))"
ncu --set full --cache-control none --clock-control none --import-source on -k regex:step_group -s 3 -c 4 -f -o gpurun_out/prof_c5 python tools/prof_c5.py > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_ncu.log
