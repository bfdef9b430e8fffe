python -m pytest tests -x -q -m gpu 2>&1 | tail -3
Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
for v in base pairtwice base pairtwice; do
  if [ $v = base ]; then L=""; else L="MRS_B200_LIB=$PWD/build_variants/lib_$v.so"; fi
  env $L python bench.py --steps 400 $Q > gpurun_out/b72_$v.json 2>>gpurun_out/b72.err; python -c "
import json
d=json.load(open('gpurun_out/b72_$v.json'))
print('$v value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))"
done
for w in c5p c2 c3; do python bench.py --steps 200 --workload $w $Q > gpurun_out/b72_$w.json 2>>gpurun_out/b72.err; python -c "
import json
d=json.load(open('gpurun_out/b72_$w.json'))
print('$w value %.3e ms/step %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))"; done
tail -2 gpurun_out/b72.err
