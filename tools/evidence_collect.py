#!/usr/bin/env python
"""Turns the output of tools/evidence.sh (gpurun_out/ev_*) into the committed files under profiles/."""
import csv
import json
import os
import re
import shutil
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(R, 'gpurun_out'), os.path.join(R, 'profiles')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r2'
sys.path.insert(0, R)


def sh(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout


for w in ('c5', 'c5_200', 'c2', 'c3', 'c4', 'c5v', 'c5p', 'm64', 'reference_arm'):
    src = os.path.join(O, 'ev_bench_%s.json' % w)
    if os.path.isfile(src) and os.path.getsize(src) > 0:
        json.load(open(src))
        shutil.copy(src, os.path.join(P, '%s_bench_%s.json' % (TAG, w)))
for w in ('c5', 'c4'):
    src = os.path.join(O, 'ev_launches_%s.csv' % w)
    if os.path.isfile(src):
        shutil.copy(src, os.path.join(P, '%s_launches_%s.csv' % (TAG, w)))
        open(os.path.join(P, '%s_launches_%s_summary.txt' % (TAG, w)), 'w').write(
            sh('python %s/tools/ncu_launch_summary.py %s' % (R, src)))

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'sm__cycles_active.max', 'sm__cycles_active.min', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']


def raw_summary(rep, dst):
    text = sh('ncu -i %s --page raw --csv' % rep)
    rows = list(csv.reader(text.splitlines()))
    if len(rows) < 3:
        return None
    h = rows[0]
    cols = [i for i, c in enumerate(h) if c in WANT or re.search(r'smsp__average_warps_issue_stalled_.*_per_issue_active', c)]
    with open(dst, 'w', newline='') as f:
        wr = csv.writer(f)
        for r in rows:
            wr.writerow([r[i] for i in cols])
    i_r, i_w = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
    scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
    single = [r for r in rows[2:] if float(r[h.index('gpu__time_duration.sum')]) < 60]    # single-step launches
    tr = [float(r[i_r]) * scale[rows[1][i_r]] + float(r[i_w]) * scale[rows[1][i_w]] for r in single]
    return sum(tr) / len(tr) if tr else None


traffic = {}
cold = raw_summary(os.path.join(O, 'ev_full_cold.ncu-rep'), os.path.join(P, '%s_ncu_full_step_group_c5.csv' % TAG))
warm = raw_summary(os.path.join(O, 'ev_full_warm.ncu-rep'), os.path.join(P, '%s_ncu_warm_cache_c5.csv' % TAG))
if cold:
    traffic['c5'] = cold
if warm:
    traffic['c5_warm'] = warm
import bench
traffic['kernel_sources_sha256'] = bench.kernel_sources_hash()
traffic['note'] = ('dram__bytes_read.sum + dram__bytes_write.sum per single-step launch of step_group_kernel<6,8,7,1,0>. c5: ncu --set '
                   'full with cold L2 (profiles/%s_ncu_full_step_group_c5.csv; reads = state + actions = the algorithmic reads; '
                   'stores still resident in the 126 MB L2 at kernel end are not counted). c5_warm: --cache-control none '
                   '(profiles/%s_ncu_warm_cache_c5.csv): what a rollout step moves through DRAM (actions in, write-back out; the '
                   'state hits L2).' % (TAG, TAG))
json.dump(traffic, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)
src = os.path.join(O, 'ev_src.csv')
open(src, 'w').write(sh('ncu -i %s --page source --csv --print-source cuda,sass --kernel-name regex:step_group --launch-skip 1 '
                        '--launch-count 1' % os.path.join(O, 'ev_full_warm.ncu-rep')))
open(os.path.join(P, '%s_ncu_source_hotlines_c5.txt' % TAG), 'w').write(
    sh('python %s/tools/ncu_source_hotlines.py %s 30' % (R, src)))
if os.path.isfile(os.path.join(O, 'ev_latency.txt')):
    shutil.copy(os.path.join(O, 'ev_latency.txt'), os.path.join(P, '%s_latency_small_batches.txt' % TAG))
for name in ('ev_soak.txt', 'ev_gputest.log'):
    if os.path.isfile(os.path.join(O, name)):
        tail = open(os.path.join(O, name)).read().strip().splitlines()[-3:]
        open(os.path.join(P, '%s_%s' % (TAG, name[3:])), 'w').write('\n'.join(tail) + '\n')
print(json.dumps(traffic, indent=1))
