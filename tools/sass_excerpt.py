"""Per-kernel count of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = legacy
mma.sync.  Usage: python tools/sass_excerpt.py [libb2n.so] > profiles/r2_sass_excerpt.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "ssl_cr_histo_b200/libb2n.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "SYNCS", "RED", "ATOM"]
cur, rows, first = None, collections.OrderedDict(), {}
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not m:
        continue
    op = m.group(1)
    rows[cur]["_total"] += 1
    for p in pats:
        if op.startswith(p):
            rows[cur][p] += 1
            first.setdefault((cur, p), ln.strip()[:110])
arch = re.findall(r"arch = (sm_\w+)", sass)
print("# SASS evidence, `cuobjdump -sass %s`\n" % lib)
print("Architectures in the fat binary: %s.  %d kernels.\n" % (sorted(set(arch)), len(rows)))
print("| kernel (demangled prefix) | SASS instr | UTCHMMA (tcgen05.mma) | LDTM (tcgen05.ld) | UTMALDG (TMA load) | UBLKCP (bulk copy) | HMMA (legacy) |")
print("|---|---:|---:|---:|---:|---:|---:|")
dem = subprocess.run(["c++filt"], input="\n".join(rows), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for (k, c), d in zip(rows.items(), dem):
    name = re.sub(r"\(.*", "", d).replace("void b2n::", "").replace("b2n::", "")
    print("| `%s` | %d | %d | %d | %d | %d | %d |" % (name[:70], c["_total"], c["UTCHMMA"] + c["UTCQMMA"] + c["UTCMMA"],
                                                 c["LDTM"], c["UTMALDG"], c["UBLKCP"], c["HMMA"]))
    tot.update(c)
print("| **total** | %d | %d | %d | %d | %d | %d |" % (tot["_total"], tot["UTCHMMA"] + tot["UTCQMMA"] + tot["UTCMMA"],
                                                   tot["LDTM"], tot["UTMALDG"], tot["UBLKCP"], tot["HMMA"]))
print("\nFirst occurrence of each mnemonic in the layer-3/4 forward kernel "
      "(`conv_igemm_kernel<256,128,2,true,false,false,-1,4>`):\n\n```")
for (k, p), ln in first.items():
    if "Li256ELi128ELi2ELb1ELb0ELb0ELin1ELi4" in k and p in ("UTCHMMA", "LDTM", "UTMALDG", "SYNCS"):
        print(ln)
print("```")
