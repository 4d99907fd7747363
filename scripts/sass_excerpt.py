#!/usr/bin/env python
"""Writes profiles/r2_sass_excerpt.txt: tcgen05 / TMEM / TMA instruction counts of the shipped libb200unet.so (whole library
and per tensor-core kernel) plus a short disassembly excerpt, from `cuobjdump -sass`.  Runs without a GPU.

  python scripts/sass_excerpt.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unet-pytorch_b200", "libb200unet.so")
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS.PHASECHK", "SYNCS.ARRIVE", "LDGSTS"]


def strip(ln):
    return re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip()


def main(out_path):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout.split("\n")
    tot, per, cur = collections.Counter(), collections.OrderedDict(), None
    for ln in sass:
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if "/*" not in ln:
            continue
        for op in OPS:
            if op in ln:
                tot[op] += 1
                if cur:
                    per.setdefault(cur, collections.Counter())[op] += 1
    out = ["# cuobjdump -sass unet-pytorch_b200/libb200unet.so -- the library the tests and bench.py load (scripts/sass_excerpt.py)",
           "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = TMA tensor load / store,",
           "# UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc/dealloc, SYNCS.* = mbarrier, LDGSTS = cp.async", "",
           "## whole library"]
    out += [f"{op:16s} {tot[op]}" for op in OPS]
    out += [f"UTCHMMA.2CTA     {sum(1 for ln in sass if 'UTCHMMA.2CTA' in ln)}   (cta_group::2 kernels)", "",
            "## per tensor-core kernel  (conv_igemm_kernel<N tile, taps, stacked M tiles, taps per B stage, A slots, B slots, "
            "mask-stream depth, UP = decoder conv with interpolation warps>)"]
    for fn, c in per.items():
        if c["UTCHMMA"] == 0:
            continue
        dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(CUtensorMap_st.*", "", dem)
        out.append(f"{dem:62s} UTCHMMA {c['UTCHMMA']:3d}  LDTM {c['LDTM']:2d}  UTMALDG {c['UTMALDG']:2d}  UTMASTG {c['UTMASTG']:2d}  "
                   f"UTCBAR {c['UTCBAR']:2d}")
    key = "conv_igemm_kernelILi256ELi9ELi1ELi1ELi2ELi4ELi0ELi1"
    start = next((i for i, ln in enumerate(sass) if key in ln and "Function" in ln), None)
    if start is not None:
        def excerpt(op, before, after):
            k = next(k for k in range(start, len(sass)) if op in sass[k])
            return [s for s in (strip(ln) for ln in sass[k - before:k + after]) if s.strip()]
        out += ["", "## excerpt: conv_igemm_kernel<256, 9, 1, 1, 2, 4, 0, 1> (decoder conv: one halo box, nine taps = nine descriptor offsets)",
                "# MMA issue (one elected thread)"] + excerpt("UTCHMMA", 8, 24)
        out += ["# epilogue: accumulators TMEM -> registers"] + excerpt("LDTM", 3, 6)
        out += ["# epilogue: swizzled staging tile -> global through TMA"] + excerpt("UTMASTG", 3, 4)
        out += ["# producer: TMA box load"] + excerpt("UTMALDG", 3, 4)
    with open(out_path, "w") as f:
        f.write("\n".join(out) + "\n")
    print(out_path, len(out), "lines")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_excerpt.txt"))
