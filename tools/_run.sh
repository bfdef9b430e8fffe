# scratch command file for `gpurun -- 'bash tools/_run.sh'` (overwritten freely during development);
# this version is the round-end check: GPU tests, smoke, the default bench line and the reference arm
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err
python -c "
import json
d=json.load(open('gpurun_out/final_bench.json'))
print('value %.3e ms/step %.4f frac %.3f traffic %s | flushed %.4f | many %.3e | e2e %.3e | cpu %.3e (%s cores) | launches %d | clocks %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['traffic'],d['l2_flushed']['ms_per_step_median'],d['step_many']['value'],d['e2e']['value'],d['cpu_baseline']['value'],d['cpu_baseline']['cores'],d['gpu_launches'],d['clocks']))"
python bench.py --impl reference --steps 20 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('reference arm value %.3e cores %s'%(d['value'], d['cpu_baseline']['cores']))"
