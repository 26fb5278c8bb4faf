"""Developer tool: per-source-line instruction / stall-sample shares of one kernel in an ncu report.
usage: ncu_lines.py report.ncu-rep <mangled-name-prefix> <cubin-name> <source-file> [min_pct]"""
import csv, re, subprocess, sys, io, os, tempfile
rep, sym, cub, srcfile = sys.argv[1:5]
minpct = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
so = '/root/repo/mddatasetbuilder_b200/libmdb_b200.so'
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
dis = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split('\n')
start = [i for i, l in enumerate(dis) if l.startswith('.text.' + sym)][0]
print(dis[start][:120])
cur = None; amap = {}
for l in dis[start + 1:]:
    if l.startswith('\t.section'): break
    m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: amap[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:100])
h = rows[1]
ia = h.index('Address'); ii = h.index('Instructions Executed'); it = h.index('Thread Instructions Executed'); isamp = h.index('# Samples')
base = None; agg = {}; tot = 0
for r in rows[2:]:
    if len(r) <= ii: continue
    a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    if base is None: base = a
    ln = amap.get(a - base)
    n = int(r[ii] or 0); th = int(r[it] or 0); sm = int(r[isamp] or 0)
    d = agg.setdefault(ln, [0, 0, 0]); d[0] += n; d[1] += th; d[2] += sm; tot += n
tots = sum(v[2] for v in agg.values())
print('total warp-inst', tot, 'samples', tots)
src = open(srcfile).read().split('\n')
byfile = {}
for ln, v in agg.items():
    f = ln[0] if ln else '?'
    d = byfile.setdefault(f, [0, 0]); d[0] += v[0]; d[1] += v[2]
for f, v in byfile.items(): print('  file %-32s inst %5.1f%% samples %5.1f%%' % (f, 100 * v[0] / tot, 100 * v[1] / tots))
for ln, v in sorted(agg.items(), key=lambda kv: (kv[0] or ('', 0))):
    if v[0] > tot * minpct / 100 or v[2] > tots * minpct / 100:
        t = src[ln[1] - 1].strip()[:90] if ln and ln[0] == os.path.basename(srcfile) else str(ln)
        print('%-5s inst %5.2f%% simt %4.1f samp %5.2f%% | %s' % (ln[1] if ln else '?', 100 * v[0] / tot, v[1] / max(v[0], 1), 100 * v[2] / tots, t))
