"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: python tools/srcprof.py dump.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], newline='')))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
agg = collections.OrderedDict()
cur_file = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr[2:], r[2:]))
    def f(k):
        try: return float(d.get(k, 0) or 0)
        except ValueError: return 0.0
    key = (cur_file, ln, r[1].strip()[:110])
    a = agg.setdefault(key, [0, 0, 0, 0])
    a[0] += f("# Samples"); a[1] += f("Instructions Executed"); a[2] += f("Thread Instructions Executed"); a[3] += f("stall_long_sb")
tot = [sum(a[i] for a in agg.values()) for i in range(4)]
print("total samples %d, warp instr %.3g, thread instr %.3g, avg threads %.1f" % (tot[0], tot[1], tot[2], tot[2] / max(tot[1], 1)))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% ins thr %4.1f lsb %4.0f%% | %s:%d %s" % (100 * a[0] / tot[0], 100 * a[1] / tot[1], a[2] / max(a[1], 1), 100 * a[3] / max(a[0], 1), key[0], key[1], key[2]))
