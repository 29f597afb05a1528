"""profiles/sass_rNN_summary.md: tcgen05 / TMA instruction counts per kernel of the built library (cuobjdump -sass; no GPU needed).
usage: python tools/sass_summary.py r02"""
import collections
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = subprocess.run(["cuobjdump", "-sass", "pytorch3dunet_b200/libb200unet.so"], capture_output=True, text=True).stdout
pats = {"UTCHMMA": r"\bUTCHMMA\b", "UTMALDG": r"\bUTMALDG", "LDTM": r"\bLDTM\b", "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS\.", "UTMASTG": r"\bUTMASTG",
        "ELECT": r"\bELECT\b", "R2UR": r"\bR2UR"}
cur, cnt = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                cnt[cur][k] += 1
dem = subprocess.run(["c++filt"], input="\n".join(cnt.keys()), capture_output=True, text=True).stdout.splitlines()
rows = [(d.split("(")[0].replace("void ", ""), c) for d, c in zip(dem, cnt.values()) if c["UTCHMMA"] or c["UTMALDG"]]
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
md = [f"# SASS of the shipped library (round {tag[1:]}): tcgen05 / TMA instruction counts per kernel\n",
      "`cuobjdump -sass pytorch3dunet_b200/libb200unet.so` (sm_100a), counted per function: `UTCHMMA` = tcgen05.mma, `UTMALDG` = TMA tile loads",
      "(cp.async.bulk.tensor), `LDTM` = tcgen05.ld, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops.  No `UTMASTG`: epilogues store with 128-bit `STG`",
      "(output rows are 32-128 B, written straight from registers).  Regenerate: `python tools/sass_summary.py`.\n",
      "| kernel | UTCHMMA | UTMALDG | LDTM | UTCBAR | SYNCS | ELECT | R2UR |", "|---|---:|---:|---:|---:|---:|---:|---:|"]
for n, c in rows:
    md.append(f"| `{n[:90]}` | {c['UTCHMMA']} | {c['UTMALDG']} | {c['LDTM']} | {c['UTCBAR']} | {c['SYNCS']} | {c['ELECT']} | {c['R2UR']} |")
md.append(f"| **total** | {tot['UTCHMMA']} | {tot['UTMALDG']} | {tot['LDTM']} | {tot['UTCBAR']} | {tot['SYNCS']} | {tot['ELECT']} | {tot['R2UR']} |")
open(f"profiles/sass_{tag}_summary.md", "w").write("\n".join(md) + "\n")
print("\n".join(md))
