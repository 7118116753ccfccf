#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: per-launch time, DRAM/L2 traffic, pipe utilisation, top stalls."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts%'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex%'),
        ('sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'issue%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs'),
        ('lts__t_sectors_srcunit_tex_op_read.sum', 'l2_rd_sectors'), ('lts__t_sectors_srcunit_tex_op_write.sum', 'l2_wr_sectors'),
        ('lts__t_sector_hit_rate.pct', 'l2hit%')]
stall = [h for h in hdr if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
for d in data:
    print(d[idx['Kernel Name']][:60])
    print('   ' + '  '.join('%s=%s%s' % (n, d[idx[k]][:9], units[idx[k]][:6]) for k, n in want if k in idx))
    tot = sum(float(d[idx[h]] or 0) for h in stall) or 1.0
    top = sorted(stall, key=lambda h: -float(d[idx[h]] or 0))[:6]
    print('   stalls: ' + ', '.join('%s %.0f%%' % (h.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * float(d[idx[h]]) / tot) for h in top))
