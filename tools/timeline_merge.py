#!/usr/bin/env python3
"""Merged view of a CF_TC_TRACE timeline (see timeline_report.py): one line per event, a column per role."""
import sys
rows = [tuple(int(x) for x in l.split()) for l in open(sys.argv[1]) if l.strip()]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 60)
t0 = min(r[2] for r in rows)
names = {10: 'x start', 27: 'x issued', 30: 'H ready', 31: 'gates issued', 40: 'rh ready', 41: 'cand issued', 49: 'R enter', 50: 'G landed',
         51: 'rh handed', 52: 'u done', 53: 'C landed', 55: 'H handed', 56: 'C left', 70: 'landed', 71: 'slab'}
roles = ['issA', 'issB', 'epiA', 'epiB', 'conv', 'prod']
for t, rg, tag, step in sorted((r[2] - t0, r[0], r[1] % 1000, r[1] // 1000) for r in rows):
    if lo <= t < hi:
        n = names.get(tag) or ('tma%d' % (tag - 80) if tag >= 80 else 'full%d' % (tag - 60) if tag >= 60 else 'iss%d' % (tag - 20))
        print("%6d  %s%-5s %s (step %d)" % (t, '        ' * rg, roles[rg], n, step))
