Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
for w in c5v c5p; do
for v in base v6p7 v7p7; do
  if [ $v = base ]; then L=""; else L="MRS_B200_LIB=$PWD/build_variants/lib_$v.so"; fi
  env $L python bench.py --steps 400 --workload $w $Q > gpurun_out/b78_${w}_$v.json 2>>gpurun_out/b78.err; python -c "
import json
d=json.load(open('gpurun_out/b78_${w}_$v.json'))
print('$w $v value %.3e ms/step %.4f frac %.3f | flushed ms %.4f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['step_many'] and '%.3e'%d['step_many']['value']))"
done; done
tail -2 gpurun_out/b78.err
