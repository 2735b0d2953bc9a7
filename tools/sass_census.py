"""SASS opcode census of libmellow_b200.so: per kernel family, how many tcgen05 MMAs (UTC*MMA), TMA loads / stores
(UTMALDG / UTMASTG), bulk prefetches (UBLKPF), TMEM loads (LDTM), legacy tensor-core MMAs (HMMA) and cp.async copies
(LDGSTS) the compiled code contains (B200_PROFILING.md, "What proves a Blackwell-native kernel").

    python tools/sass_census.py [--out profiles/r2_sass_census.txt]
"""
import argparse, collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(ROOT, "mellow_b200", "csrc", "libmellow_b200.so"))
ap.add_argument("--out", default="")
args = ap.parse_args()
sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "UTMAPF", "UTMACCTL", "LDTM", "STTM", "HMMA", "LDGSTS", "PRMT", "FFMA"]
fam = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        dem = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"^void ", "", dem).replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        base = re.split(r"[<(]", name)[0].replace("mb::", "")
        cur = fam.setdefault(base, {"variants": 0, "ops": collections.Counter()})
        cur["variants"] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                cur["ops"][w] += 1
lines = [f"# SASS opcode census of {os.path.relpath(args.lib, ROOT)} (cuobjdump -sass, sm_100a); counts summed over the template variants",
         f"{'kernel':38s} {'variants':>8s} " + " ".join(f"{w:>8s}" for w in WATCH)]
tot = collections.Counter()
for base, d in fam.items():
    lines.append(f"{base:38s} {d['variants']:8d} " + " ".join(f"{d['ops'][w]:8d}" for w in WATCH))
    tot.update(d["ops"])
lines.append(f"{'TOTAL':38s} {sum(d['variants'] for d in fam.values()):8d} " + " ".join(f"{tot[w]:8d}" for w in WATCH))
text = "\n".join(lines)
print(text)
if args.out:
    with open(args.out, "w") as f:
        f.write(text + "\n")
