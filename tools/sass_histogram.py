#!/usr/bin/env python
"""SASS opcode histogram of the headline kernel (step_group_kernel<set_speeds, N=8, 7 warps, baked, single step>) and of
the out-of-line contact path it calls, from the object file the build produced:

    python tools/sass_histogram.py > profiles/r2_sass_histogram_step_group_c5.txt

Static counts (one per instruction in the binary, not per execution); the dynamic count per launch is ncu's
smsp__inst_executed.sum in profiles/r2_ncu_*."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(REPO, 'mrs-gym_b200', 'build', 'mode6.o')
KERNEL = '_ZN3mrs17step_group_kernelILi6ELi8ELi7ELb1ELb0EEEv9MrsConfigNS_7DerivedE10MrsBuffersNS_8StepArgsE'


def main():
    out = subprocess.run(['cuobjdump', '-sass', '-fun', KERNEL, OBJ], capture_output=True, text=True).stdout
    sections, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = 'kernel ' + m.group(1)
            sections[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)', line)
        if m and cur:
            op = m.group(2).rstrip(';')
            sections[cur][op] += 1
            # the dump appends the functions the kernel calls behind its own (unpredicated) EXIT
            if op == 'EXIT' and not m.group(1) and cur.startswith('kernel'):
                cur = 'called from it: chunk_step_contact (contact path, out of line) and compiler helpers (division, sqrt)'
                sections[cur] = collections.Counter()
    if not sections:
        sys.exit('no SASS found in %s (build first)' % OBJ)
    # cuobjdump lists the kernel and the functions it calls in one dump: the first section is the kernel
    for name, h in sections.items():
        total = sum(h.values())
        print('== %s\n   %d instructions' % (name, total))
        groups = collections.OrderedDict([
            ('FP32 arithmetic (FFMA FMUL FADD FMNMX FSEL FSETP ...)', lambda o: o[0] == 'F' and not o.startswith('FLO')),
            ('MUFU (rcp / rsq / ex2 / sqrt)', lambda o: o.startswith('MUFU')),
            ('integer / address (IADD3 IMAD LEA LOP3 SHF MOV ...)', lambda o: re.match(r'(IADD|IMAD|LEA|LOP3|SHF|MOV|VIADD|ISETP|SEL|UIADD|ULEA|UMOV|UISETP|ULOP|USHF|UIMAD|PLOP|P2R|R2P|PRMT|HFMA2|I2F|F2I|POPC|FLO|BREV)', o) is not None),
            ('LDGSTS (cp.async global -> shared)', lambda o: o.startswith('LDGSTS')),
            ('LDS / STS (shared memory)', lambda o: o.startswith('LDS') or o.startswith('STS')),
            ('LDG', lambda o: o.startswith('LDG') and not o.startswith('LDGSTS') and not o.startswith('LDGDEPBAR')),
            ('STG.E.128 / STG.E.EF.128', lambda o: o.startswith('STG') and '128' in o),
            ('STG other', lambda o: o.startswith('STG') and '128' not in o),
            ('LDC / LDCU (constant bank)', lambda o: o.startswith('LDC')),
            ('LDL / STL (local memory)', lambda o: o.startswith('LDL') or o.startswith('STL')),
            ('SHFL', lambda o: o.startswith('SHFL')),
            ('VOTE / VOTEU / REDUX / MATCH / ELECT', lambda o: re.match(r'(VOTE|REDUX|MATCH|ELECT)', o) is not None),
            ('ATOMS / ATOMG / RED', lambda o: re.match(r'(ATOM|RED\b|REDG)', o) is not None),
            ('branches / convergence (BRA BSSY BSYNC CALL RET EXIT WARPSYNC)', lambda o: re.match(r'(BRA|BSSY|BSYNC|CALL|RET|EXIT|WARPSYNC|BREAK|JMP)', o) is not None),
            ('barriers / fences / PDL (BAR MEMBAR DEPBAR LDGDEPBAR PREEXIT ACQBULK CCTL ERRBAR)', lambda o: re.match(r'(BAR|MEMBAR|DEPBAR|LDGDEPBAR|PREEXIT|ACQBULK|CCTL|ERRBAR|CGAERRBAR)', o) is not None),
            ('S2R / S2UR / CS2R', lambda o: re.match(r'(S2R|S2UR|CS2R)', o) is not None),
        ])
        seen = set()
        for label, pred in groups.items():
            ops = {o: n for o, n in h.items() if pred(o) and o not in seen}
            seen |= set(ops)
            if ops:
                print('   %5d  %s' % (sum(ops.values()), label))
        rest = {o: n for o, n in h.items() if o not in seen}
        if rest:
            print('   %5d  other: %s' % (sum(rest.values()), ', '.join('%s %d' % kv for kv in sorted(rest.items(), key=lambda kv: -kv[1])[:12])))
        print('   top opcodes: ' + ', '.join('%s %d' % kv for kv in h.most_common(28)))


if __name__ == '__main__':
    main()
