"""Developer tool (no GPU needed): per-kernel SASS mnemonic counts of libmdb_b200.so -> profiles/rNN_sass_summary.md.
usage: python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "mddatasetbuilder_b200/libmdb_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
rows, cur = [], None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = {"name": m.group(1), "n": 0, "mma": 0, "ldtm": 0, "tma": 0, "bar": 0, "dfma": 0, "hmma": 0}
        rows.append(cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if cur is None or not m:
        continue
    op = m.group(1)
    cur["n"] += 1
    if op.startswith("UTC") and "MMA" in op: cur["mma"] += 1
    if op.startswith("LDTM"): cur["ldtm"] += 1
    if op.startswith("UTMALDG"): cur["tma"] += 1
    if op.startswith("UTCBAR") or op.startswith("SYNCS"): cur["bar"] += 1
    if op.startswith("DFMA"): cur["dfma"] += 1
    if op.startswith("HMMA") or op.startswith("DMMA"): cur["hmma"] += 1
keep = [r for r in rows if re.search(r"k_bonds|k_tql|k_tridiag|k_assign|k_composition|k_kpp_persistent|k_feature_rows|k_mb_update", r["name"])]
print("# SASS summary of `libmdb_b200.so` (cuobjdump -sass, sm_100a), round 2, final code\n")
print("Mnemonics that prove the tcgen05 / TMEM / TMA path (`/opt/skills/guides/B200_PROFILING.md`), per kernel; built by "
      "`mddatasetbuilder_b200/csrc/build.sh`; this table by `tools/sass_summary.py`.\n")
print("| kernel | SASS instructions | UTCHMMA / UTC*MMA | LDTM (tcgen05.ld) | UTMALDG (TMA load) | UTCBAR / SYNCS | DFMA | HMMA/DMMA |\n|---|---|---|---|---|---|---|---|")
for r in keep:
    print("| `%s` | %d | %d | %d | %d | %d | %d | %d |" % (r["name"][:72], r["n"], r["mma"], r["ldtm"], r["tma"], r["bar"], r["dfma"], r["hmma"]))
print("\nTotals over the library (%d kernels): %d SASS instructions, %d UTC*MMA, %d LDTM, %d UTMALDG." %
      (len(rows), sum(r["n"] for r in rows), sum(r["mma"] for r in rows), sum(r["ldtm"] for r in rows), sum(r["tma"] for r in rows)))
