"""Per-source-line dynamic instruction counts from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_by_line.py export.csv [kernel-substring] [top]"""
import csv, collections, sys
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
kern = cur_file = hdr = None
per_line = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name':
        kern = r[1]; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 9 or r[0] == '' or want not in (kern or ""):
        continue
    try:
        inst, samp, tinst = int(r[7]), int(r[6]), int(r[8])
    except ValueError:
        continue
    k = (cur_file, int(r[0]))
    per_line[k][0] += inst; per_line[k][1] += samp; per_line[k][2] += tinst
tot = sum(v[0] for v in per_line.values()); tots = sum(v[1] for v in per_line.values())
print("kernel filter %r: %d warp instructions, %d samples" % (want, tot, tots))
srcs = {}
def text(f, ln):
    import os
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "wolkenbase_b200", "csrc", f)
        srcs[f] = open(p).read().split('\n') if os.path.exists(p) else None
    return srcs[f][ln - 1].strip()[:88] if srcs[f] and ln <= len(srcs[f]) else ""
acc = 0
for (f, ln), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    acc += v[0]
    print("%5.2f%% inst %5.2f%% smp  lanes %4.1f  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / max(1, tots), v[2] / max(1, v[0]), f, ln, text(f, ln)))
print("top %d lines cover %.1f%%" % (top, 100 * acc / tot))
