"""Developer tool: ncu launch-list CSV (--metrics gpu__time_duration.sum) -> the markdown table kept under profiles/.
usage: ncu_launches.py launches.csv "<command line that was profiled>" > profiles/rNN_launches.md"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if 'Kernel Name' in r)
ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
own, other = {}, {}
for r in rows[rows.index(hdr) + 1:]:
    if len(r) != len(hdr) or r[im] != 'gpu__time_duration.sum': continue
    name = r[ik].split('(')[0].strip()
    d = own if (name.startswith('k_') or ' k_' in name) else other
    e = d.setdefault(name, [0, 0.0]); e[0] += 1; e[1] += float(r[iv].replace(',', '')) / 1e3   # ns -> us
tot = sum(v[1] for v in own.values())
print('# launch list\n')
print('`%s`' % sys.argv[2])
print('(run after the same command exited 0 without ncu; cold-cache, serialised: compare SHARES with bench.py\'s live `kernels` shares)\n')
print('| kernel | launches | total ms | share of own kernels | avg us |\n|---|---|---|---|---|')
for k, (n, us) in sorted(own.items(), key=lambda kv: -kv[1][1]):
    print('| `%s` | %d | %.3f | %.3f | %.1f |' % (k, n, us / 1e3, us / tot, us / n))
print('\nOther kernels in the same process (torch generators / memset, not part of the timed path): ' +
      ', '.join('`%s` x%d %.2f ms' % (k[:40], n, us / 1e3) for k, (n, us) in sorted(other.items(), key=lambda kv: -kv[1][1])[:6]))
