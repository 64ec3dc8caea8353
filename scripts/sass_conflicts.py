"""Static estimate of FP32 operand-fetch stalls in a kernel's SASS.

Empirical rule read off the ncu source page of the Debye kernel on B200
(profiles/): an FFMA whose A and B source registers (first two sources) are
fetched from the register file in the same cycle and have the same parity
(register index mod 2) takes one extra dispatch cycle; an operand carried over
by the .reuse cache of the previous instruction does not count.

usage: python scripts/sass_conflicts.py <lib.so> <function-substring>
"""
import collections
import re
import subprocess
import sys


def functions(path):
    txt = subprocess.check_output(['cuobjdump', '-sass', path], text=True)
    parts = re.split(r'\n\s*Function : ', txt)
    return {p.split('\n', 1)[0].strip(): p for p in parts[1:]}


def analyze(body):
    ins = []
    for line in body.split('\n'):
        m = re.search(r'/\*([0-9a-f]+)\*/\s+(.*?);', line)
        if m:
            ins.append(m.group(2).strip())
    prev = {}
    stats = collections.Counter()
    for s in ins:
        s = re.sub(r'^@!?U?P\d+\s+', '', s)
        opc = s.split()[0].split('.')[0]
        if opc not in ('FFMA', 'FMUL', 'FADD'):
            prev = {}
            continue
        srcs = s.split(None, 1)[1].split(',')[1:]
        regs, new = {}, {}
        for slot, o in enumerate(srcs):
            m = re.search(r'(?<![A-Z])R(\d+)(\.reuse)?', o)
            if not m:
                continue
            r = int(m.group(1))
            if prev.get(slot) != r:
                regs[slot] = r
            if m.group(2):
                new[slot] = r
        prev = new
        stats[opc] += 1
        if 0 in regs and 1 in regs and regs[0] % 2 == regs[1] % 2 and regs[0] != regs[1]:
            stats[opc + '_ab_conflict'] += 1
    return stats


if __name__ == '__main__':
    for name, body in functions(sys.argv[1]).items():
        if sys.argv[2] in name:
            st = analyze(body)
            n = st['FFMA'] + st['FMUL'] + st['FADD']
            c = st['FFMA_ab_conflict'] + st['FMUL_ab_conflict'] + st['FADD_ab_conflict']
            print(name[:70], dict(st), 'fp32 instr', n, 'A/B same-parity', c,
                  'pred cyc/instr %.3f' % ((n + c) / max(n, 1)))
