"""Per-kernel SASS opcode evidence of the in-tree library (no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
Counts the mnemonics B200_PROFILING.md names as proof of the Blackwell paths: UBLKCP (1-D bulk async copy = TMA engine),
UTMALDG (TMA tensor loads), UTC*MMA / LDTM / STTM (tcgen05 + TMEM), DMMA (FP64 tensor-core MMA), LDGSTS (cp.async),
SYNCS (mbarrier), REDUX / MATCH (warp-level primitives of the radix planner)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

so = Path(__file__).resolve().parent.parent / "polars_ols_b200" / "libb200ols.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
KEYS = ["UBLKCP", "UTMALDG", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "DMMA", "DFMA", "LDGSTS", "SYNCS", "MATCH", "REDUX", "ATOMS", "STG", "LDG", "LDS"]
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[kern][k] += 1
        counts[kern]["_total"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {so.name}: {len(counts)} kernels, sm_100a SASS (cuobjdump -sass); columns = static instruction counts")
print("# " + " ".join(f"{k:>8}" for k in KEYS) + "    total  kernel")
tot = collections.Counter()
for (k, c), name in zip(counts.items(), demangle):
    name = re.sub(r"\(.*$", "", name).replace("b200::", "")
    print("  " + " ".join(f"{c[x]:>8}" for x in KEYS) + f" {c['_total']:>8}  {name}")
    tot.update(c)
print("# " + " ".join(f"{tot[x]:>8}" for x in KEYS) + f" {tot['_total']:>8}  ALL KERNELS")
print("# tcgen05 (UTC*MMA / LDTM / STTM) and TMA tensor loads (UTMALDG): none in the library — see DESIGN.md 4.1 / 4.11 (no f64 kind; the tf32 probe tools/tf32_gram_probe.cu, which does contain UTCHMMA / LDTM, measured why not for k = 16 f32)")
