"""Static evidence for profiles/: per-kernel SASS opcode counts of the shipped libtemd.so (cuobjdump -sass) and the
ptxas register / spill / shared-memory lines of a verbose rebuild.  Runs on the CPU box (no GPU needed).
    python tools/sass_report.py          -> profiles/r02_sass_opcodes.txt, profiles/r02_ptxas.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytemdiags_b200 import build as B  # noqa: E402

KEYS = ('DMMA', 'UTMALDG', 'SYNCS', 'USETMAXREG', 'LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL', 'DFMA', 'DMUL', 'DADD', 'BAR', 'SHFL',
        'UTCHMMA', 'UTCIMMA', 'LDTM', 'HMMA', 'IMMA')


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', B.LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and cur:
            op = m.group(1).split('.')[0]
            counts[cur][op] += 1
            counts[cur]['_total'] += 1
    names = demangle(list(counts))
    with open(os.path.join(ROOT, 'profiles', 'r02_sass_opcodes.txt'), 'w') as f:
        f.write('cuobjdump -sass pytemdiags_b200/libtemd.so : static opcode counts per kernel (sm_100a)\n')
        f.write('kernel | total | ' + ' | '.join(KEYS) + '\n')
        tot = collections.Counter()
        for k, c in counts.items():
            short = re.sub(r'\(.*', '', names.get(k, k)).replace('temd::', '')
            f.write('%s | %d | %s\n' % (short, c['_total'], ' | '.join(str(c[x]) for x in KEYS)))
            tot.update(c)
        f.write('ALL KERNELS | %d | %s\n' % (tot['_total'], ' | '.join(str(tot[x]) for x in KEYS)))
        f.write('\nDMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64); UTMALDG = TMA bulk tensor load; SYNCS = mbarrier ops;\n'
                'USETMAXREG = setmaxnreg; UTC*MMA / LDTM (tcgen05) absent by design: tcgen05 has no f64 kind.\n')
    # ptxas -v of every translation unit
    lines = []
    for src in B.SOURCES:
        cmd = ['nvcc'] + B.NVCC_FLAGS + ['-Xptxas', '-v', '-c', os.path.join(B.CSRC, src), '-o', '/dev/null']
        out = subprocess.run(cmd, capture_output=True, text=True).stderr
        fn = None
        for line in out.split('\n'):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                fn = m.group(1)
            elif 'bytes stack frame' in line and fn:
                stack = line.strip()
            elif line.strip().startswith('ptxas info    : Used') and fn:
                lines.append((src, fn, stack, line.strip().replace('ptxas info    : ', '')))
                fn = None
    names = demangle([l[1] for l in lines])
    with open(os.path.join(ROOT, 'profiles', 'r02_ptxas.txt'), 'w') as f:
        f.write('nvcc %s -Xptxas -v : registers / spills / shared memory per kernel\n' % ' '.join(B.NVCC_FLAGS))
        for src, fn, stack, used in lines:
            short = re.sub(r'\(.*', '', names.get(fn, fn)).replace('temd::', '')
            f.write('%-18s %-52s %s ; %s\n' % (src, short, used, stack))
    print('wrote profiles/r02_sass_opcodes.txt (%d kernels), profiles/r02_ptxas.txt (%d entries)' % (len(counts), len(lines)))


if __name__ == '__main__':
    main()
