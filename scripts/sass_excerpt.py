"""Regenerate profiles/r02_sass_tcgen05_tma_excerpt.txt: per kernel, the counts and first occurrences of the Blackwell
instructions that prove which hardware paths the shipped objects use (cuobjdump -sass of lcgan_b200/csrc/_build/*.o)."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT",
        "ELECT", "FHFMA", "LDGSTS")
FIRST = ("UTCHMMA", "UTMALDG", "FHFMA")


def main():
    out = ["# cuobjdump -sass lcgan_b200/csrc/_build/{conv_tc,warp,resample}.o (sm_100a), end of round 2 (scripts/sass_excerpt.py):",
           "# UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit ->",
           "# mbarrier, SYNCS = mbarrier ops, UTCATOMSWS = TMEM alloc, FHFMA = fma.rn.f32.bf16 (mixed-precision FMA), LDGSTS = cp.async", ""]
    for obj in ("conv_tc.o", "warp.o", "resample.o"):
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "lcgan_b200", "csrc", "_build", obj)],
                              capture_output=True, text=True).stdout
        kernels, cur = collections.OrderedDict(), None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "")
                cur = kernels.setdefault(name, [])
            elif cur is not None and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
                cur.append(line.strip())
        out.append(f"#### {obj}")
        for name, lines in kernels.items():
            cnt = collections.Counter()
            for l in lines:
                for w in WANT:
                    if re.search(r"\b" + re.escape(w), l):
                        cnt[w] += 1
            if not cnt:
                continue
            out.append(f"== {name}: " + ", ".join(f"{k} x{v}" for k, v in sorted(cnt.items())))
            for w in FIRST:
                hits = [l for l in lines if re.search(r"\b" + w, l)][:3]
                out += ["   " + re.sub(r"\s+", " ", h) for h in hits]
            out.append("")
    path = os.path.join(ROOT, "profiles", "r02_sass_tcgen05_tma_excerpt.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print(path, len(out), "lines")


if __name__ == "__main__":
    main()
