python -m pytest tests -x -q -m gpu 2>&1 | tail -3
Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
for pd in 0 16 8 32 16; do
MRS_B200_POOL_DIV=$pd python bench.py --steps 400 $Q > gpurun_out/b71_$pd.json 2>>gpurun_out/b71.err; python -c "
import json
d=json.load(open('gpurun_out/b71_$pd.json'))
print('pool 1/$pd value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))"
done
tail -2 gpurun_out/b71.err
timeout 200 python tools/soak.py c5 5000 2>&1 | tail -1
