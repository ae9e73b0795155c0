"""Summarise an ncu report (run where ncu is installed): python profiles/ncu_summary.py <report.ncu-rep> [out.md] [--traffic ENTRY]
--traffic ENTRY also records dram__bytes_read.sum + dram__bytes_write.sum of every kernel of the report in profiles/field_traffic.json
under the C-ABI entry point name ENTRY (e.g. avc_eval_occupancy_grid): that file is where bench.py takes `roofline.traffic` from."""
import csv, io, json, os, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__t_bytes.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
STALLS = 'smsp__average_warps_issue_stalled_'
def to_bytes(val, unit):
    v = float(val.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)
def main():
    entry = None
    if '--traffic' in sys.argv:
        i = sys.argv.index('--traffic'); entry = sys.argv[i + 1]; del sys.argv[i:i + 2]
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        out.append('## %s' % d.get('Kernel Name', ('', '?'))[1])
        for k in KEYS:
            if k in d: out.append('- %s = %s %s' % (k, d[k][1], d[k][0]))
        st = sorted(((float(v[1] or 0), h[len(STALLS):].replace('_per_issue_active.ratio', '')) for h, v in d.items() if h.startswith(STALLS) and h.endswith('per_issue_active.ratio')), reverse=True)[:6]
        out.append('- top stalls (warps per issue): ' + ', '.join('%s %.2f' % (n, x) for x, n in st))
        if entry and 'dram__bytes_read.sum' in d:
            name = d['Kernel Name'][1].replace('void ', '').replace('<unnamed>::', '').split('(')[0].split('<')[0]
            rec = {'kernel': name, 'entry': entry, 'dram_bytes': int(to_bytes(d['dram__bytes_read.sum'][1], d['dram__bytes_read.sum'][0]) + to_bytes(d['dram__bytes_write.sum'][1], d['dram__bytes_write.sum'][0])),
                   'dram_read_bytes': int(to_bytes(d['dram__bytes_read.sum'][1], d['dram__bytes_read.sum'][0])), 'dram_write_bytes': int(to_bytes(d['dram__bytes_write.sum'][1], d['dram__bytes_write.sum'][0])),
                   'source': 'ncu --set full, ' + os.path.basename(rep)}
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'field_traffic.json')
            recs = [r for r in (json.load(open(path)) if os.path.exists(path) else []) if not (r.get('kernel') == name and r.get('entry') == entry)]
            recs.append(rec); json.dump(recs, open(path, 'w'), indent=1)
    txt = '\n'.join(out)
    if len(sys.argv) > 2: open(sys.argv[2], 'w').write(txt + '\n')
    print(txt)
if __name__ == '__main__':
    main()
