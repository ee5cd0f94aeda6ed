"""Prints the counters the profiles/ summaries quote from an `ncu --page raw --csv` export (one block per launch) and, with
--source FILE (an `ncu --page source --csv` export), the stall samples aggregated over 1 KB blocks of SASS.
Usage: python tools/ncu_summary.py raw.csv [--source source.csv]"""
import csv, sys, collections

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__waves_per_multiprocessor', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    stall = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('== ' + d.get('Kernel Name', '?'))
        for w in WANT:
            if w in d:
                print(f'{w:90s} {d[w]}')
        st = sorted(((float(d[h]), h.split('issue_stalled_')[1].split('_per_')[0]) for h in stall if d[h] not in ('', 'n/a')),
                    reverse=True)
        print('stalls per issue: ' + ' '.join(f'{n}={v:.2f}' for v, n in st[:10]))


def source(path, block=64):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0] != 'Address']

    def I(r, k):
        try:
            return int(r[ix[k]] or 0)
        except ValueError:
            return 0
    keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(I(r, '# Samples') for r in data)
    print(f'-- stall samples by {block}-instruction block ({tot} samples, {len(data)} instructions)')
    blk = collections.OrderedDict()
    for n, r in enumerate(data):
        b = blk.setdefault(n // block, collections.Counter())
        b['n'] += I(r, '# Samples')
        b['exec'] += I(r, 'Instructions Executed')
        for s in keys:
            b[s] += I(r, s)
    for k, b in blk.items():
        if b['n'] > tot * 0.004:
            top = sorted(((b[s], s[6:]) for s in keys), reverse=True)[:4]
            print(f"{k * block * 16:#7x} {100 * b['n'] / tot:5.1f}% exec {b['exec']:>10} " +
                  ' '.join(f'{n}={100 * v / max(1, b["n"]):.0f}%' for v, n in top))
    allk = collections.Counter()
    for r in data:
        for s in keys:
            allk[s[6:]] += I(r, s)
    print('total: ' + ' '.join(f'{n}={100 * v / tot:.1f}%' for n, v in allk.most_common(10)))


if __name__ == '__main__':
    raw(sys.argv[1])
    if '--source' in sys.argv:
        source(sys.argv[sys.argv.index('--source') + 1])
