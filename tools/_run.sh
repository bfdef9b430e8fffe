python -m pytest tests -x -q -m gpu 2>&1 | tail -3
Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
for w in c4; do
python bench.py --steps 200 --workload $w $Q > gpurun_out/b70_$w.json 2>>gpurun_out/b70.err; python -c "
import json
d=json.load(open('gpurun_out/b70_$w.json'))
print('$w value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | launches %d'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['gpu_launches']))"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/b70_launches_c4.csv python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu --clock-seconds 0 --e2e-steps 2 > gpurun_out/b70_l.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/b70_launches_c4.csv > gpurun_out/b70_sum.txt; head -5 gpurun_out/b70_sum.txt
timeout 200 python tools/soak.py c4 2000 2>&1 | tail -2
