"""Per-kernel SASS mnemonic counts that prove the Blackwell-native paths (run where cuobjdump is installed; no GPU needed):
    python profiles/sass_summary.py [lib.so] > profiles/r2_sass_summary.md
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load (cp.async.bulk.tensor), UBLKCP = bulk copy
(cp.async.bulk), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA would be the legacy mma.sync path (none expected)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'avatarcap_b200', 'libavatarcap_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'HMMA', 'LDG', 'STG', 'RED', 'ATOM', 'ST.E.STRONG.SYS', 'LD.E.STRONG.SYS', 'MEMBAR.SC.SYS', 'MEMBAR.ALL.SYS']
rows = []
cur = None; cnt = None; total = 0
for line in txt.splitlines():
    m = re.match(r'\s+Function : (\S+)', line)
    if m:
        if cur: rows.append((cur, cnt, total))
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip(); cnt = collections.Counter(); total = 0
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        op = m.group(1); total += 1
        for k in KEYS:
            if op == k or op.startswith(k + '.'): cnt[k] += 1
        if '.SYS' in op: cnt['SYS'] += 1
if cur: rows.append((cur, cnt, total))
print('# SASS evidence per kernel of %s (sm_100a)\n' % os.path.basename(lib))
print('| kernel | instructions | UTC*MMA (of which .2CTA) | LDTM / STTM | UTMALDG (TMA tensor) | UBLKCP (bulk copy) | UTCBAR | SYNCS (mbarrier) | HMMA | system-scope ld/st/membar |')
print('|---|---|---|---|---|---|---|---|---|---|')
for name, c, total in sorted(rows, key=lambda r: -r[2]):
    short = re.sub(r'\(anonymous namespace\)::', '', name).split('(')[0].replace('void ', '')
    sysn = c['SYS']
    print('| `%s` | %d | %d (%d) | %d / %d | %d | %d | %d | %d | %d | %d |' % (short[:90], total, c['UTCHMMA'], c['UTCHMMA.2CTA'], c['LDTM'], c['STTM'], c['UTMALDG'], c['UBLKCP'],
                                                                    c['UTCBAR'], c['SYNCS'], c['HMMA'], sysn))
