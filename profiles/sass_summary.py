"""Per-kernel counts of the Blackwell-specific SASS mnemonics in lib/libdbs_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
HMMA (warp-level mma.sync), ENL2.256 (32-byte global loads / stores).  usage: python profiles/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "diffusion-boltzmann-sampler_b200", "lib", "libdbs_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
MN = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "UTCATOMSWS", "HMMA", "LDSM", "MUFU", "POPC", "REDUX", "IMAD.WIDE", "ENL2.256"]
counts, cur, n_instr = collections.OrderedDict(), None, collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        n_instr[cur] += 1
        for k in MN:
            if re.search(r"\b" + re.escape(k) + r"\b", line) or (k + ".") in line:
                counts[cur][k] += 1
print(f"# {os.path.relpath(lib, ROOT)}: {len(counts)} kernels (sm_100a), cuobjdump -sass; columns = occurrences in the kernel's SASS")
print("# kernel | instructions | " + " | ".join(MN))
tot = collections.Counter()
for fn, c in counts.items():
    name = re.sub(r"dbs::\(anonymous namespace\)::|void |dbs::", "", demangle(fn))
    name = re.sub(r"\(CUtensorMap_st.*", "", name)[:110]
    print(f"{name} | {n_instr[fn]} | " + " | ".join(str(c[k]) for k in MN))
    tot.update(c)
print("# total | " + str(sum(n_instr.values())) + " | " + " | ".join(str(tot[k]) for k in MN))
