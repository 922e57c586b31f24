#!/usr/bin/env python
"""profiles/r2_sass_excerpts.txt: SASS mnemonic counts per kernel of the built library + the K2 MMA-issuer loop (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
LIB = os.path.join(ROOT, 'dhr_b200', 'lib', 'libdhr_b200.so')
WANT = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'STTM', 'UTMALDG', 'UTMALDG.2D.2CTA', 'UTMALDG.2D.MULTICAST', 'UTMAPF', 'UBLKCP', 'UTCBAR',
        'UTCBAR.2CTA.MULTICAST', 'UTCATOMSWS', 'SYNCS', 'HMMA', 'FHFMA', 'ATOMG', 'ATOMS', 'REDUX', 'CREDUX', 'LDS', 'STS', 'LDG', 'STG']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
    out = ['# SASS mnemonic counts per kernel of dhr_b200/lib/libdhr_b200.so (cuobjdump -sass, sm_100a; tools/sass_excerpts.py).  tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2),',
           '# tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG (.2CTA / .MULTICAST forms), cp.async.bulk -> UBLKCP, tcgen05.commit -> UTCBAR,',
           '# tcgen05.alloc -> UTCATOMSWS, mbarrier -> SYNCS, fma.rn.f32.f16 -> FHFMA.  No HMMA (legacy mma.sync) anywhere.']
    cur, counts, total, lines, bodies = None, None, 0, [], {}

    def flush():
        if cur is None:
            return
        name = re.sub(r'\(.*$', '', demangle(cur))
        parts = ['%s=%d' % (k, counts[k]) for k in WANT if counts.get(k)]
        lines.append('%-90s total %5d  %s' % (name[:90], total, '  '.join(parts)))

    for l in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', l)
        if m:
            flush()
            cur, counts, total = m.group(1), collections.Counter(), 0
            bodies[cur] = []
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m and cur:
            op = m.group(1)
            total += 1
            bodies[cur].append(re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', re.sub(r'^\s*/\*[0-9a-f]+\*/\s*', '', l)).rstrip())
            base = op.split('.')[0]
            for k in WANT:
                if op == k or (('.' not in k) and base == k and op not in WANT) or (('.' in k) and op.startswith(k)):
                    counts[k] += 1
                    break
    flush()
    out += lines
    for fn, title in (('dense_tile_ts_kernel', 'K2 cta_group::1'), ('dense_tile_ts2_kernel', 'K2 cta_group::2')):
        for name, body in bodies.items():
            if fn + 'E' in name:
                idx = [i for i, b in enumerate(body) if 'UTCHMMA' in b]
                if not idx:
                    continue
                lo, hi = idx[0], idx[-1]
                while lo > 0 and 'SYNCS.PHASECHK' not in body[lo]:
                    lo -= 1
                hi2 = hi
                while hi2 < len(body) - 1 and 'UTCBAR' not in body[hi2]:
                    hi2 += 1
                out += ['', '# %s: the MMA-issuer loop of one corpus stage -- wait on the stage-full barrier, 4 x UTCHMMA with additive descriptors, commit (%d instructions)'
                        % (title, hi2 - lo + 1)]
                out += ['    ' + b for b in body[lo:hi2 + 1]]
    open(os.path.join(ROOT, 'profiles', 'r2_sass_excerpts.txt'), 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out[-45:]))


if __name__ == '__main__':
    sys.exit(main())
