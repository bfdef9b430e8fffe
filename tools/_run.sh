python -m pytest tests -x -q -m gpu 2>&1 | tail -3
Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
for w in c5 c5p c3; do
python bench.py --steps 200 --workload $w $Q > gpurun_out/b67_$w.json 2>>gpurun_out/b67.err; python -c "
import json
d=json.load(open('gpurun_out/b67_$w.json'))
print('$w value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))"
done
