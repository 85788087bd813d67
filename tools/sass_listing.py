"""Write profiles/<name>: which kernels of libupnerf_b200.so use the Blackwell tensor / TMA instructions,
with counts and excerpts, from `cuobjdump -sass` (run in the build container; no GPU needed).

    python tools/sass_listing.py [--out profiles/r2_sass_tcgen05.txt]
"""
import argparse
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCATOMSWS", "SYNCS",
             "UBLKCP", "HMMA", "FFMA"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r2_sass_tcgen05.txt"))
    ap.add_argument("--excerpt", type=int, default=6, help="lines of context around the first UTCHMMA of a kernel")
    args = ap.parse_args()
    so = ROOT / "upnerf_b200" / "lib" / "libupnerf_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        if cur is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            kernels[cur].append(line.rstrip())
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    out = [f"# cuobjdump -sass {so.relative_to(ROOT)}  (sm_100a)", "# per kernel: instruction count and the Blackwell "
           "tensor-core / TMA / mbarrier mnemonics it contains", ""]
    rows = []
    for name, lines in kernels.items():
        cnt = {m: sum(1 for l in lines if re.search(r"\b" + m + r"[.\s]", l)) for m in MNEMONICS}
        rows.append((name, len(lines), cnt, lines))
    for name, n, cnt, lines in rows:
        if not (cnt["UTCHMMA"] or cnt["UTMALDG"] or cnt["LDTM"]):
            continue
        d = demangle(name)
        out.append(f"== {d[:160]}")
        out.append(f"   {n} SASS instructions: " + ", ".join(f"{m} x{c}" for m, c in cnt.items() if c))
        first = next((i for i, l in enumerate(lines) if "UTCHMMA" in l), None)
        if first is not None:
            lo, hi = max(0, first - args.excerpt), min(len(lines), first + args.excerpt + 1)
            out += ["   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.strip()) for l in lines[lo:hi]]
        for tag in ("LDTM", "UTMALDG", "UTMASTG"):
            l = next((l for l in lines if tag in l), None)
            if l:
                out.append("   first " + tag + ": " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.strip()))
        out.append("")
    simt = [(demangle(n)[:100], c) for n, c, cnt, _ in rows if not (cnt["UTCHMMA"] or cnt["UTMALDG"] or cnt["LDTM"])]
    out.append(f"# {len(simt)} other kernels (SIMT: streaming / per-ray work), instruction counts:")
    out += [f"   {c:6d}  {n}" for n, c in simt]
    Path(args.out).write_text("\n".join(out) + "\n")
    print(f"wrote {args.out}: {sum(1 for r in rows if r[2]['UTCHMMA'])} kernels with UTCHMMA")


if __name__ == "__main__":
    main()
