"""Regenerates profiles/sass_plan_kernel_*.txt from the in-tree libfhp.so (cuobjdump; no GPU needed).

    python profiles/make_sass_extracts.py

One file per kernel variant the evidence cites: resource usage, instruction counts by mnemonic, the memory / tensor-copy /
barrier mnemonics in full, and the first lines of the listing.  The full listing is one cuobjdump command away (header
of each file)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quadrupedal_foothold_planner_b200", "libfhp.so")
T = "_ZN3fhp11plan_kernelILb{cen}ELi{stage}ELb{d}ELb{h}ELb{s}ELi{spec}ELb{one}EEEvNS_8PlanArgsENS_6GeoDevE"
VARIANTS = {  # file tag -> template arguments <WANT_CEN, STAGE, WANT_DEF, HEADING, STREAM, SPEC, ONEPAR>
    "mask": dict(cen=0, stage=2, d=0, h=0, s=0, spec=1, one=1),        # the headline kernel (C3): yaml geometry at 0.02 m
    "mask_001m": dict(cen=0, stage=2, d=0, h=0, s=0, spec=2, one=1),   # the same at 0.01 m (C5): cp.async elevation slots
    "mask_multi": dict(cen=0, stage=2, d=0, h=0, s=0, spec=1, one=0),  # parameter table in global memory (sweeps)
    "mask_generic": dict(cen=0, stage=2, d=0, h=0, s=0, spec=0, one=0),
    "tma": dict(cen=0, stage=1, d=0, h=0, s=0, spec=0, one=0),         # 2-D TMA staging path
    "stream": dict(cen=0, stage=2, d=0, h=0, s=1, spec=1, one=1),      # fhp_plan_stream: rows into mapped host memory
    "heading": dict(cen=0, stage=2, d=0, h=1, s=0, spec=1, one=0),     # rotated rectangle (C4)
}
FULL = re.compile(r"^(LDG|STG|LDS|STS|LDL|STL|LDGSTS|UTMALDG|UBLKCP|SYNCS|BAR|SHFL|VOTE|DADD|DMUL|DFMA|LDGDEPBAR|DEPBAR|ATOM|RED)")


def main():
    for tag, args in VARIANTS.items():
        fun = T.format(**args)
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", "-fun", fun, LIB], capture_output=True, text=True).stdout
        ins = [m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)]
        if not ins:
            raise SystemExit(f"{tag}: {fun} not found in {LIB}")
        by = collections.Counter(i.split(".")[0] for i in ins)
        full = collections.Counter(i for i in ins if FULL.match(i))
        usage = [l.strip() for l in res.splitlines() if "REG:" in l]
        listing = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", l)]
        with open(os.path.join(ROOT, "profiles", f"sass_plan_kernel_{tag}.txt"), "w") as f:
            f.write(f"# cuobjdump -sass -fun '{fun}' libfhp.so  (sm_100a; profiles/make_sass_extracts.py)\n")
            f.write("# resource usage:\n  " + (usage[0] if usage else "?") + "\n")
            f.write(f"# {len(ins)} instructions; counts by mnemonic (whole function incl. out-of-line callees):\n")
            for k, v in by.most_common(40):
                f.write(f"{v:7d} {k}\n")
            f.write("# selected mnemonics (full):\n")
            for k, v in full.most_common():
                f.write(f"{v:7d} {k}\n")
            f.write("# first 60 instructions of the listing:\n")
            for l in listing[:60]:
                f.write(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() + "\n")
        print(tag, len(ins), usage[0] if usage else "?")


if __name__ == "__main__":
    main()
