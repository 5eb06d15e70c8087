"""Per-function aggregation of an ncu source-page CSV (developer tool): instructions executed, lane
efficiency and stall samples by enclosing device function. nvdisasm -g gives only the innermost
inlined line, so one-line vector helpers (svcm_device.cuh:30-70) are attributed to the last larger
function seen in address order.

usage: python tools/ncu_funcs.py <source-page.csv> <nvdisasm -g -c output> <mangled-kernel-substring>
"""
import collections, csv, os, re, sys
csv.field_size_limit(10**9)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_csv, sass_txt, kname = sys.argv[1:4]

def functions(path):
    out = []      # (first line, name)
    pat = re.compile(r'^(?:template<[^>]*>\s*)?(?:SVCM_HD|SVCM_D|__global__|static|inline|__device__)[^;{=]*?\b([A-Za-z_]\w*)\s*\(')
    for n, line in enumerate(open(path), 1):
        m = pat.match(line)
        if m and m.group(1) not in ('__launch_bounds__', 'SVCM_LB_LIGHT'):
            out.append((n, m.group(1)))
    return out
files = {}
for f in ('svcm_device.cuh', 'svcm_kernels.cu', 'svcm_kutil.cuh', 'svcm_bvh.h'):
    files[f] = functions(os.path.join(ROOT, 'smallvcm_b200', 'csrc', f))
def func_of(f, line):
    best = None
    for n, name in files.get(f, []):
        if n <= line: best = name
        else: break
    return best or f
# kernels are found by their own signature lines (k_xxx( at line start)
for n, line in enumerate(open(os.path.join(ROOT, 'smallvcm_b200', 'csrc', 'svcm_kernels.cu')), 1):
    m = re.match(r'^(k_\w+)\(', line)
    if m: files['svcm_kernels.cu'].append((n, m.group(1)))
files['svcm_kernels.cu'].sort()

addr2 = {}; cur = None; inside = False
for line in open(sass_txt, errors='replace'):
    if line.startswith('//---') and '.text.' in line:
        inside = kname in line; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m and cur: addr2[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = None; first = None; agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]; last = '?'
for r in rows:
    if r and r[0] == 'Address': h = r; continue
    if not h or len(r) < len(h): continue
    try: a = int(r[0], 16) if r[0].startswith('0x') else int(r[0])
    except ValueError: continue
    if first is None: first = a
    f, line = addr2.get(a - first, ('?', 0))
    helper = (f == 'svcm_device.cuh' and line < 72) or f.endswith('.hpp') or f.endswith('.h') and f != 'svcm_bvh.h' or f == '?'
    name = last if helper else func_of(f, line)
    last = name
    ie = int(float(r[h.index('Instructions Executed')] or 0)); te = int(float(r[h.index('Thread Instructions Executed')] or 0))
    ss = int(float(r[h.index('# Samples')] or 0))
    for t, v in zip((agg[name], tot), ((ie, te, ss),) * 2):
        t[0] += v[0]; t[1] += v[1]; t[2] += v[2]
print('total warp-inst %d  lanes %.1f/32  samples %d' % (tot[0], tot[1] / max(1, tot[0]), tot[2]))
for name, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] * 200 < tot[0]: continue
    print('%-28s inst %5.1f%%  samples %5.1f%%  lanes %4.1f' % (name, 100.0 * v[0] / tot[0], 100.0 * v[2] / max(1, tot[2]), v[1] / max(1, v[0])))
