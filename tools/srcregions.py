"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by code region of hvb_core.cuh / hvb_kernels.cuh.
usage: python tools/srcregions.py dump.csv   (line ranges below match the sources of the commit the dump was taken at)"""
import csv, sys, collections
REG = [("hvb_core.cuh", 202, 223, "load_x32"), ("hvb_core.cuh", 231, 250, "hash/edge_slot"), ("hvb_core.cuh", 254, 313, "ortho_direction/dot"),
       ("hvb_core.cuh", 327, 380, "best/verify64"), ("hvb_core.cuh", 382, 403, "make_filter"), ("hvb_core.cuh", 409, 517, "row_range32/row_try32"),
       ("hvb_core.cuh", 518, 561, "row_range(fp64)"), ("hvb_core.cuh", 562, 666, "scan_points"), ("hvb_core.cuh", 667, 678, "settle_stage"),
       ("hvb_core.cuh", 679, 711, "query: planes"), ("hvb_core.cuh", 712, 792, "query: stage setup"), ("hvb_core.cuh", 793, 896, "query: row loop"),
       ("hvb_core.cuh", 897, 943, "vertex_insert"), ("hvb_core.cuh", 944, 1003, "edge_register"), ("hvb_core.cuh", 1004, 1081, "commit_vertex"),
       ("hvb_core.cuh", 1082, 1216, "ray_setup/ray_result"), ("hvb_kernels.cuh", 418, 473, "k_walk loop"),
       ("hvb_coop.cuh", 40, 84, "coop: row geometry"), ("hvb_coop.cuh", 85, 139, "coop: next_row (task fetch)"), ("hvb_coop.cuh", 140, 206, "coop: scan chunk + survivors"),
       ("hvb_coop.cuh", 207, 326, "coop: stage setup / settle"), ("hvb_coop.cuh", 327, 462, "commit_vertex_warp"), ("hvb_coop.cuh", 463, 600, "k_walk_coop loop")]
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
