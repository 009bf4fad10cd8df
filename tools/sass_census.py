"""Per-kernel SASS census of libhdrnet_b200.so: which Blackwell-specific instructions each kernel
actually contains (cuobjdump -sass), as a markdown table.
    python tools/sass_census.py > profiles/r02_sass_census.md"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hdrnet_b200", "lib", "libhdrnet_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = [("UTCHMMA", "tcgen05.mma"), ("LDTM", "tcgen05.ld"), ("STTM", "tcgen05.st"), ("UTCBAR", "tcgen05.commit"),
        ("UBLKCP", "cp.async.bulk (TMA 1-D)"), ("UTMALDG", "TMA tensor"), ("SYNCS", "mbarrier"), ("FFMA2", "fma.rn.f32x2"),
        ("FMUL2", "mul.rn.f32x2"), ("TLD", "tex1Dfetch"), ("LDGSTS", "cp.async"), ("LDS", "ld.shared"), ("STS", "st.shared"),
        ("MUFU", "sfu"), ("BAR.SYNC", "bar.sync")]
cur, counts, total = None, collections.OrderedDict(), {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1); total[cur] += 1
        for k, _ in KEYS:
            if op == k or op.startswith(k + ".") or (k == "BAR.SYNC" and op.startswith("BAR.SYNC")):
                counts[cur][k] += 1
print("# SASS census of `libhdrnet_b200.so` (sm_100a), one row per kernel\n")
print("`cuobjdump -sass`, instruction counts per kernel (static, not executed counts).  " + "; ".join(f"{k} = {v}" for k, v in KEYS) + ".\n")
print("| kernel | SASS instr | " + " | ".join(k for k, _ in KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
agg = collections.Counter()
for fn in counts:
    name = re.sub(r"hdrnet_b200::", "", demangle(fn))
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    print(f"| `{name[:100]}` | {total[fn]} | " + " | ".join(str(counts[fn][k]) if counts[fn][k] else "" for k, _ in KEYS) + " |")
    agg.update(counts[fn])
print(f"| **all {len(counts)} kernels** | {sum(total.values())} | " + " | ".join(str(agg[k]) for k, _ in KEYS) + " |")
