"""Per-kernel SASS instruction counts of libarcb200.so (cuobjdump -sass): which kernels carry FP64
tensor-core (DMMA), per-thread cp.async (LDGSTS), TMA-unit bulk copies (UBLKCP / UTMALDG), mbarrier
(SYNCS, ARRIVES) instructions and local-memory spills (STL / LDL).
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "arc-alkali-rydberg-calculator_b200", "libarcb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["DMMA", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "ARRIVES", "BAR", "ATOMG", "RED", "STL", "LDL"]
kern, counts = None, collections.OrderedDict()
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)[.\s;]", l)
    if m and kern:
        counts[kern]["instr"] += 1
        for k in keys:
            if m.group(2).startswith(k):
                counts[kern][k] += 1
print("SASS instruction counts per kernel of libarcb200.so (cuobjdump -sass, sm_100a; tools/sass_summary.py).")
print("UBLKCP = cp.async.bulk (TMA unit), SYNCS = mbarrier ops, ARRIVES = cp.async.mbarrier.arrive,")
print("DMMA = mma.sync.m8n8k4.f64, LDGSTS = per-thread cp.async, STL/LDL = local-memory spills.\n")
print("%-40s %7s" % ("kernel", "instr") + "".join(" %7s" % k for k in keys))
for k, c in counts.items():
    if c["instr"]:
        print("%-40s %7d" % (k[:40], c["instr"]) + "".join(" %7d" % c[x] for x in keys))
