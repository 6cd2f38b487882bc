"""Census of the Blackwell-specific SASS in the built library: per kernel, how many tcgen05 (UTC*MMA), TMEM (LDTM/STTM),
TMA (UTMALDG/UTMASTG), mbarrier (SYNCS), cluster (UCGABAR, ST.ASYNC / STAS, MAPA) instructions it holds.
  python tools/sass_census.py [lib.so] > profiles/rNN_sass_census.txt"""
import collections, re, subprocess, sys

PAT = re.compile(r'\b(UTC[A-Z]*MMA[.\w]*|UTCBAR[.\w]*|UTCCP[.\w]*|UTCATOMSWS[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|'
                 r'SYNCS[.\w]*|UCGABAR[.\w]*|STAS[.\w]*|MAPA[.\w]*|CCTL[.\w]*|MUFU\.EX2|HMMA[.\w]*|DFMA|DMUL|DADD)\b')


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else 'e2e_multi_view_matching_b200/libmvm_b200.so'
    txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    fn = None
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r'\(anonymous namespace\)::', '', fn).split('(')[0]
            per[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = PAT.search(line)
        if m:
            per[fn][m.group(1)] += 1
    print('# cuobjdump -sass %s : Blackwell-specific instruction census per kernel (static counts)' % lib)
    for fn, c in per.items():
        keep = {k: v for k, v in c.items() if not k.startswith(('DFMA', 'DMUL', 'DADD', 'CCTL')) or v}
        if not keep:
            continue
        print('%s' % fn)
        for k, v in sorted(keep.items()):
            print('    %-34s %d' % (k, v))


if __name__ == '__main__':
    main()
