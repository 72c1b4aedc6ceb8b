"""ncu .ncu-rep -> markdown + json summary of the metrics the roofline needs.  usage: ncu_summary.py rep out_prefix title cmd"""
import csv, io, json, subprocess, sys
rep, out, title, cmd = sys.argv[1:5]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed']
def num(s):
    try: return float(s.replace(',', ''))
    except Exception: return s
md = [f'# {title}', '', f'Command: `{cmd}`', '']
js = {}
seen = {}
for d in data:
    name = d[hdr.index('Kernel Name')]
    key = name.split('(')[0].replace('void ', '').strip()
    seen[key] = seen.get(key, 0) + 1
    if seen[key] > 1: continue
    md += [f'## `{name[:110]}`', '', '| metric | value | unit |', '|---|---:|---|']
    rec = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w); v = num(d[i]); rec[w] = {'value': v, 'unit': units[i]}
            md.append(f'| {w} | {d[i]} | {units[i]} |')
    st = sorted(((num(d[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and d[i]), reverse=True)[:4]
    md += ['', 'Top issue-stall reasons (warps per issue-active cycle): ' + ', '.join(f'{h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")} {v:.2f}' for v, h in st), '']
    def byt(k):
        r = rec.get(k)
        if not r: return None
        return r['value'] * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r['unit'], 1)
    rd, wr = byt('dram__bytes_read.sum'), byt('dram__bytes_write.sum')
    rec['traffic_bytes'] = (rd + wr) if rd is not None and wr is not None else None
    js[key] = rec
open(out + '.md', 'w').write('\n'.join(md) + '\n')
json.dump(js, open(out + '.json', 'w'), indent=1)
print('\n'.join(md[:60]))
