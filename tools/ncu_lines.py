"""Joins an ncu per-SASS-instruction CSV (ncu -i X.ncu-rep --page source --csv --kernel-name ...)
with nvdisasm -g line info of the cubin to aggregate executed instructions and stall samples per
CUDA source line / inlined function. Developer tool for reading profiles without a GUI.

usage: python tools/ncu_lines.py <source-page.csv> <nvdisasm -g -c output> <mangled-kernel-substring> [topN]
"""
import csv, re, sys, collections
csv.field_size_limit(10**9)
src_csv, sass_txt, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# --- address -> (file, line, inlined-at chain) from nvdisasm
addr2line = {}
cur = None
inside = False
for line in open(sass_txt, errors='replace'):
    if line.startswith('//---') and '.text.' in line:
        inside = kname in line
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), 'inlined' in m.group(3))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m and cur:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
h = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
first = None
for r in rows:
    if r and r[0] == 'Address': h = r; first = None; continue
    if not h or len(r) < len(h) or not r[0].startswith('0x') and not r[0].isdigit(): continue
    a = int(r[0], 16) if r[0].startswith('0x') else int(r[0])
    if first is None: first = a
    off = a - first
    ie = int(float(r[h.index('Instructions Executed')] or 0)); te = int(float(r[h.index('Thread Instructions Executed')] or 0))
    ss = int(float(r[h.index('# Samples')] or 0))
    key = addr2line.get(off, (('?', 0, False), ''))[0]
    agg[key][0] += ie; agg[key][1] += te; agg[key][2] += ss
    tot[0] += ie; tot[1] += te; tot[2] += ss
print('total warp-inst %d thread-inst %d samples %d  (lane eff %.1f/32)' % (tot[0], tot[1], tot[2], tot[1] / max(1, tot[0])))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print('%-20s:%-5d %s inst %5.1f%%  samples %5.1f%%  lanes %.1f' % (key[0], key[1], 'I' if key[2] else ' ', 100.0 * v[0] / max(1, tot[0]), 100.0 * v[2] / max(1, tot[2]), v[1] / max(1, v[0])))
