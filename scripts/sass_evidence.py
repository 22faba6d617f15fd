"""Counts the Blackwell-specific SASS mnemonics per kernel of libddn_b200.so (cuobjdump -sass, no GPU needed) and writes
profiles/<tag>_sass_evidence.md: UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor loads, LDTM = tcgen05.ld (TMEM -> registers),
UTCBAR = tcgen05.commit -> mbarrier, UTCATOMSWS = TMEM alloc/dealloc, SYNCS.* = mbarrier ops, REDG...F32x4 = red.global.add.v4.f32.

    python scripts/sass_evidence.py r1
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch-dense-correspondence_b200", "libddn_b200.so")
PAT = re.compile(r"\b(UTCHMMA\.2CTA|UTCHMMA|UTMALDG\.\dD\.2CTA|UTMALDG(?:\.\dD)?|LDTM(?:\.x\d+)?|UTCBAR\.2CTA\.MULTICAST|UTCBAR|UTCATOMSWS|"
                 r"SYNCS(?:\.[A-Z0-9]+)*|REDG\.E\.ADD\.F32x4|REDG\.E\.ADD\.F64|HMMA|FFMA)\b")
# longest prefixes first: the .2CTA forms are the cta_group::2 (CTA-pair) variants
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG.4D.2CTA", "UTMALDG.2D.2CTA", "UTMALDG", "LDTM", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "UTCATOMSWS",
        "SYNCS", "REDG.E.ADD.F32x4", "REDG.E.ADD.F64", "HMMA", "FFMA"]


def main(tag):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    demangled = dict(zip(re.findall(r"Function : (\S+)", sass), names))
    counts = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1); counts[fn] = collections.Counter(); continue
        if fn is None:
            continue
        m = PAT.search(line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    counts[fn][k] += 1
                    break
    out = os.path.join(ROOT, "profiles", tag + "_sass_evidence.md")
    with open(out, "w") as f:
        f.write("# %s: Blackwell SASS mnemonics per kernel of libddn_b200.so (`cuobjdump -sass`, sm_100a)\n\n" % tag)
        f.write("UTCHMMA = `tcgen05.mma` (.2CTA = `cta_group::2`, the CTA-pair form); UTMALDG = TMA tensor load (.2CTA = pair form signalling the\n"
                "leader's mbarrier); LDTM = `tcgen05.ld`; UTCBAR = `tcgen05.commit` (.2CTA.MULTICAST = to both CTAs of the pair); UTCATOMSWS = TMEM\n"
                "alloc/dealloc; SYNCS = mbarrier; REDG.E.ADD.F32x4 = `red.global.add.v4.f32`; REDG.E.ADD.F64 = the BatchNorm-statistics reds;\n"
                "HMMA = legacy `mma.sync` (none expected).\n"
                "Static instruction counts (loops are not unrolled into them), kernels without any tensor/TMA instruction listed last.\n\n")
        f.write("| kernel | " + " | ".join(KEYS) + " |\n|---|" + "---:|" * len(KEYS) + "\n")
        rows = sorted(counts.items(), key=lambda kv: (-kv[1]["UTCHMMA.2CTA"], -kv[1]["UTCHMMA"], -kv[1]["FFMA"]))
        for fn, c in rows:
            name = re.sub(r"\(.*", "", demangled.get(fn, fn)).replace("void ", "")
            f.write("| `%s` | " % name[:80] + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |\n")
    print("wrote", out, "(%d kernels)" % len(counts))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r1")
