"""One line per kernel launch of an ncu --set full report: duration, DRAM bytes, achieved GB/s against the measured HBM peak.
    python profiles/ncu_aux_summary.py <report.ncu-rep | raw.csv> [out.md]"""
import csv, io, json, os, subprocess, sys
def main():
    src = sys.argv[1]
    raw = open(src).read() if src.endswith('.csv') else subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        peak = float(json.load(open(os.path.join(root, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        peak = 6650.0
    def num(d, k, unit_scale=None):
        u, v = d.get(k, ('', '0'))
        x = float((v or '0').replace(',', ''))
        if unit_scale:
            x *= unit_scale.get(u, 1.0)
        return x
    B = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    T = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'second': 1.0, 's': 1.0, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9}
    out = ['| kernel | grid x block | time (us) | DRAM read (MB) | DRAM write (MB) | achieved GB/s | of measured HBM peak (%.0f GB/s) | L2 throughput %% | top stall |' % peak,
           '|---|---|---|---|---|---|---|---|---|']
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        name = d.get('Kernel Name', ('', '?'))[1].split('(')[0].replace('<unnamed>::', '')
        t = num(d, 'gpu__time_duration.sum', T)
        rd = num(d, 'dram__bytes_read.sum', B); wr = num(d, 'dram__bytes_write.sum', B)
        st = sorted(((float((v[1] or '0').replace(',', '')), h) for h, v in d.items() if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('per_issue_active.ratio')), reverse=True)
        top = st[0][1][len('smsp__average_warps_issue_stalled_'):].replace('_per_issue_active.ratio', '') if st else ''
        gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
        out.append('| %s | %s x %s | %.1f | %.2f | %.2f | %.0f | %.1f %% | %s | %s |' % (
            name, d.get('launch__grid_size', ('', '?'))[1], d.get('launch__block_size', ('', '?'))[1], t * 1e6, rd / 1e6, wr / 1e6, gbs, 100 * gbs / peak,
            d.get('lts__throughput.avg.pct_of_peak_sustained_elapsed', ('', '?'))[1], top))
    txt = '\n'.join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(txt + '\n')
    print(txt)
if __name__ == '__main__':
    main()
