Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
for r in 0 2 3 4 5; do
MRS_B200_PAIR_RES=$r python bench.py --steps 200 --workload c4 $Q > gpurun_out/b61_$r.json 2>>gpurun_out/b61.err; python -c "
import json
d=json.load(open('gpurun_out/b61_$r.json'))
print('res $r value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac']))"
done
tail -3 gpurun_out/b61.err
