"""One line per kernel launch of an .ncu-rep: time, throughputs, occupancy, lane efficiency, top stalls."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
stall = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_warp_active.pct')]
def g(r, k):
    return r[hdr.index(k)] if k in hdr else ''
for r in rows[2:]:
    st = sorted([(float(r[hdr.index(h)] or 0), h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', '')) for h in stall], reverse=True)[:4]
    print('%-14s %8s %s | sm%% %5s dram%% %5s l1hit %5s l2hit %5s | warps%% %5s lanes %5s issue%% %5s inst %11s regs %s | dramR %s %s dramW %s %s | %s' % (
        g(r, 'Kernel Name')[:14], g(r, 'gpu__time_duration.sum')[:8], units[hdr.index('gpu__time_duration.sum')],
        g(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed')[:5], g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[:5],
        g(r, 'l1tex__t_sector_hit_rate.pct')[:5], g(r, 'lts__t_sector_hit_rate.pct')[:5],
        g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')[:5], g(r, 'smsp__thread_inst_executed_per_inst_executed.ratio')[:5],
        g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')[:5], g(r, 'smsp__inst_executed.sum')[:11], g(r, 'launch__registers_per_thread'),
        g(r, 'dram__bytes_read.sum')[:7], units[hdr.index('dram__bytes_read.sum')], g(r, 'dram__bytes_write.sum')[:7], units[hdr.index('dram__bytes_write.sum')],
        ' '.join('%s:%.0f' % (b, a) for a, b in st)))
