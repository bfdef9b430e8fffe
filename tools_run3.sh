python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/t3.log
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $Q > gpurun_out/b3_mb6.json 2>gpurun_out/b3.err
MRS_B200_LIB=$PWD/build_variants/lib_mb4.so python bench.py $Q > gpurun_out/b3_mb4.json 2>>gpurun_out/b3.err
MRS_B200_LIB=$PWD/build_variants/lib_mb8.so python bench.py $Q > gpurun_out/b3_mb8.json 2>>gpurun_out/b3.err
for bps in 2 3 4 5; do MRS_B200_BLOCKS_PER_SM=$bps python bench.py $Q > gpurun_out/b3_mb6_bps$bps.json 2>>gpurun_out/b3.err; done
BA="--steps 20 --warmup 3 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $BA > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_group -s 30 -c 2 -o gpurun_out/prof_r1b python bench.py $BA > gpurun_out/ncu3.log 2>&1
cat gpurun_out/t3.log; tail -3 gpurun_out/b3.err
for f in gpurun_out/b3_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %.3e'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many']['value']))
"; done
