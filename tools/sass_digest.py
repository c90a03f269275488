"""SASS digest of libfgs_b200.so (no GPU needed): per kernel the count of the Blackwell mnemonics that prove the
tensor-core / TMEM / TMA paths (B200_PROFILING.md): UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st),
UTMALDG / UTMASTG (cp.async.bulk.tensor load / store), UTCBAR (tcgen05.commit), SYNCS (mbarrier), RED / ATOMG, plus
registers and static shared memory from the ELF.   usage: python tools/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fgs_nerf_b200", "libfgs_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP", "SYNCS", "REDG", "RED.", "ATOMG",
       "LDGSTS", "HMMA", "FFMA", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[name]["instr"] += 1
            for p in PAT:
                if op.startswith(p):
                    kernels[name][p] += 1
            if ".2CTA" in op and op.startswith("UTCHMMA"):
                kernels[name]["UTCHMMA.2CTA"] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a), instruction counts per kernel; demangled names shortened")
    tot = collections.Counter()
    for k, c in kernels.items():
        dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem.replace("(anonymous namespace)::", ""))
        flags = "  ".join(f"{p} {c[p]}" for p in PAT + ["UTCHMMA.2CTA"] if c[p])
        print(f"{dem}\n    instr {c['instr']}  {flags}\n    {usage.get(k, '')}")
        tot.update(c)
    print("# library totals: " + "  ".join(f"{p} {tot[p]}" for p in PAT + ["UTCHMMA.2CTA"] if tot[p]))


if __name__ == "__main__":
    main()
