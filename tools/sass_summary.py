"""Per-kernel SASS evidence of the Blackwell-native path: counts of the tcgen05 / TMEM / TMA mnemonics in the shipped
library (cuobjdump -sass; runs without a GPU).

    python tools/sass_summary.py > profiles/r02_sass_summary.json

UTCHMMA / UTCQMMA ... = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / .st (TMEM), UTMALDG = TMA
tensor load (cp.async.bulk.tensor), UBLKCP = cp.async.bulk, UTMAREDG = TMA reduce, SYNCS = mbarrier, HMMA = mma.sync
(the head_dim 32/64 attention only), MUFU.EX2 = exp2.
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mla_b200", "lib", "libmla_b200.so")
PAT = {"tcgen05_mma": r"\bUTC[A-Z]*MMA", "tcgen05_commit": r"\bUTCBAR", "tmem_ld": r"\bLDTM", "tmem_st": r"\bSTTM",
       "tma_tensor_load": r"\bUTMALDG", "bulk_copy": r"\bUBLKCP", "mbarrier": r"\bSYNCS", "mma_sync": r"\bHMMA",
       "ex2": r"MUFU\.EX2", "local_spill": r"\b(STL|LDL)\b"}


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        if cur is None or "/*" not in line:
            continue
        cur["instructions"] += 1
        for k, p in PAT.items():
            if re.search(p, line):
                cur[k] += 1
    res = {k: dict(v) for k, v in kernels.items()}
    tc = {k: v for k, v in res.items() if v.get("tcgen05_mma")}
    json.dump({"library": os.path.relpath(LIB, ROOT), "kernels_total": len(res), "kernels_with_tcgen05_mma": sorted(tc),
               "kernels": res}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
