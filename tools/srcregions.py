"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by code region of hvb_core.cuh / hvb_kernels.cuh.
usage: python tools/srcregions.py dump.csv   (line ranges below match the sources of the commit the dump was taken at)"""
import csv, sys, collections
REG = [("hvb_core.cuh", 203, 228, "load_x32"), ("hvb_core.cuh", 236, 258, "hash/edge_slot"), ("hvb_core.cuh", 259, 322, "ortho_direction/dot"),
       ("hvb_core.cuh", 323, 403, "best/verify64/tie_window"), ("hvb_core.cuh", 404, 431, "make_filter"), ("hvb_core.cuh", 432, 542, "row_range32/row_try32"),
       ("hvb_core.cuh", 543, 587, "row_range(fp64)"), ("hvb_core.cuh", 588, 691, "scan_points"), ("hvb_core.cuh", 692, 702, "settle_stage"),
       ("hvb_core.cuh", 703, 727, "query: planes"), ("hvb_core.cuh", 728, 816, "query: stage setup"), ("hvb_core.cuh", 817, 929, "query: row loop"),
       ("hvb_core.cuh", 930, 974, "vertex_insert"), ("hvb_core.cuh", 975, 1033, "edge_register"), ("hvb_core.cuh", 1034, 1104, "commit_vertex"),
       ("hvb_core.cuh", 1105, 1255, "ray_setup/ray_result"),
       ("hvb_coop.cuh", 40, 88, "coop: row geometry"), ("hvb_coop.cuh", 89, 215, "coop_scan (row tickets)"),
       ("hvb_coop.cuh", 216, 283, "pool: chunk filter + survivors"), ("hvb_coop.cuh", 284, 368, "pool: round setup + row slots"),
       ("hvb_coop.cuh", 369, 384, "pool: chunk task emission"), ("hvb_coop.cuh", 385, 396, "pool: chunk loop"),
       ("hvb_coop.cuh", 397, 520, "coop: stage setup / settle"), ("hvb_coop.cuh", 521, 657, "commit_vertex_warp"), ("hvb_coop.cuh", 658, 800, "k_walk_coop loop")]
rows = list(csv.reader(open(sys.argv[1], newline='')))
agg = collections.OrderedDict(); cur = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr[2:], r[2:]))
    def f(k):
        try: return float(d.get(k, 0) or 0)
        except ValueError: return 0.0
    name = cur
    for fn, a, b, nm in REG:
        if cur == fn and a <= ln <= b: name = nm
    a = agg.setdefault(name, [0, 0, 0, 0, 0])
    a[0] += f("# Samples"); a[1] += f("Instructions Executed"); a[2] += f("Thread Instructions Executed"); a[3] += f("stall_long_sb"); a[4] += f("stall_no_inst")
tot = [sum(a[i] for a in agg.values()) for i in range(5)]
print("total samples %d, warp instr %.3g, thread instr %.3g, avg threads %.1f" % (tot[0], tot[1], tot[2], tot[2] / max(tot[1], 1)))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%5.1f%% smp %5.1f%% warp-ins %5.1f%% thr-ins  thr/ins %4.1f  long_sb %3.0f%%  no_inst %3.0f%% | %s" % (100 * a[0] / tot[0], 100 * a[1] / tot[1], 100 * a[2] / tot[2], a[2] / max(a[1], 1), 100 * a[3] / max(a[0], 1), 100 * a[4] / max(a[0], 1), k))
