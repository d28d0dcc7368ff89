"""Per-CUDA-source-line instruction and stall totals from an ncu report captured with --import-source on.
usage: python tools/ncu_source_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kern,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); iw = hdr.index('Warp Stall Sampling (All Samples)')
agg = collections.OrderedDict()
cur = None
for r in rows[hi + 1:]:
    if len(r) <= ia: continue
    if r[0].strip().isdigit():          # a CUDA source line; SASS rows that follow belong to it
        cur = (int(r[0]), r[isrc])
        agg.setdefault(cur, [0, 0])
        continue
    if cur is None: continue
    try:
        agg[cur][0] += int(r[ia] or 0); agg[cur][1] += int(r[iw] or 0)
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"total warp-inst {tot}  stall samples {tots}")
for (ln, src), (a, w) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a/tot*100:6.2f}% inst {w/max(tots,1)*100:6.2f}% stall | {ln:5d} | {src.strip()[:120]}")
