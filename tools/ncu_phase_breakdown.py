"""Instruction / stall share of each phase of k_ht_encode from an ncu report captured with --import-source on.
usage: python tools/ncu_phase_breakdown.py report.ncu-rep [nblocks]
The phases are located by marker strings in grok_b200/csrc/ht_enc.cu (the source the report was built from must match)."""
import csv, subprocess, io, collections, sys, os
rep = sys.argv[1]
nblocks = int(sys.argv[2]) if len(sys.argv) > 2 else 49728
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:k_ht_encode',
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]; ia = hdr.index('Instructions Executed'); iw = hdr.index('Warp Stall Sampling (All Samples)')
agg = collections.Counter(); st = collections.Counter(); cur = None
for r in rows[hi + 1:]:
    if len(r) <= ia: continue
    if r[0].strip().isdigit(): cur = int(r[0]); continue
    if cur is None: continue
    try: agg[cur] += int(r[ia] or 0); st[cur] += int(r[iw] or 0)
    except ValueError: pass
tot = sum(agg.values()); ts = sum(st.values())
src = open(os.path.join(ROOT, 'grok_b200/csrc/ht_enc.cu')).read().split('\n')
def find(s):
    return [i + 1 for i, l in enumerate(src) if s in l][0]
marks = [('ring_get15', 1), ('gather', find('uint64_t gather64')), ('ms_drain32', find('template <bool FINAL>')), ('ms_drain128', find('void ms_drain128')),
         ('vlc_drain32', find('int vlc_drain32')), ('vlc_drain128', find('void vlc_drain128')), ('mel', find('struct Mel')), ('bits_put', find('void bits_put')), ('sm_exp', find('int sm_exponent')),
         ('layout', find('struct EncLayout')), ('prologue', find('k_ht_encode(const HtBlockDesc*')), ('staging', find('/* ---- stage: sample rows')),
         ('inlane', find('/* ---- code: lane = unit')), ('join_mel', find('/* ---- join: MEL events')), ('scan', find('/* ---- join: where each unit')),
         ('ms_join', find('/* ---- join: MagSgn.')), ('vlc_join', find('/* ---- join: VLC, ')), ('terminate', find('/* ---- terminate MagSgn')),
         ('end', find('/* lengths -> exclusive byte offsets'))]
for (n, a), (_, b) in zip(marks, marks[1:]):
    v = sum(c for l, c in agg.items() if a <= l < b); s = sum(c for l, c in st.items() if a <= l < b)
    print(f"{n:12s} {v/tot*100:6.2f}% inst {s/max(ts,1)*100:6.2f}% stall  {v/nblocks:8.0f} inst/block")
print('total inst/block', tot / nblocks)
