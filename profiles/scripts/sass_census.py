"""SASS census of libquark_b200.so: per kernel, instruction count and the mnemonics that prove the Blackwell-native path.
usage: python profiles/scripts/sass_census.py > profiles/r02_sass_census.md   (needs cuobjdump; no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = os.path.join(ROOT, "unified_audio_b200", "lib", "libquark_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEYS = ["UTCHMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2"]
kern, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["_n"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                kern[cur][k] += 1
print("# Round 2 - SASS census of every kernel in libquark_b200.so (`cuobjdump -sass`, sm_100a; `profiles/scripts/sass_census.py`)\n")
print("Mnemonics that prove the Blackwell-native path: `UTCHMMA` = tcgen05.mma (kind::f16), `UTCCP` = tcgen05.cp, `LDTM` / `STTM` = tcgen05.ld / st (TMEM <-> registers),\n"
      "`UTMALDG` = TMA tensor load, `UBLKCP` = cp.async.bulk, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops; `HMMA` = legacy mma.sync.\n")
print("| kernel | SASS instr | " + " | ".join(KEYS) + " |")
print("|---|---:|" + "---:|" * len(KEYS))
rows = sorted(kern.items(), key=lambda kv: (-(kv[1]["UTCHMMA"] > 0), -(kv[1]["HMMA"] > 0), -kv[1]["_n"]))
for name, c in rows:
    d = re.sub(r"\((int|bool|unsigned int)\)", "", demangle(name))
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    print(f"| `{d[:80]}` | {c['_n']} | " + " | ".join(str(c[k]) for k in KEYS) + " |")
tc = [n for n, c in kern.items() if c["UTCHMMA"]]
print(f"\n{len(kern)} kernels; {len(tc)} issue tcgen05.mma; {sum(1 for c in kern.values() if c['HMMA'])} use legacy mma.sync "
      "(the LM decode / prefill-continuation kernels, the round-1 attention kept for A/B, the cross-check LSTM).")

# ---- excerpts: every tensor-core / TMEM / TMA instruction of the tcgen05 kernels, verbatim (profiles/r02_sass_excerpts.txt)
if len(sys.argv) > 1:
    pat = re.compile(r"\b(UTCHMMA|UTCCP|LDTM|STTM|UTMALDG|UBLKCP|UTCBAR|UTCATOMSWS)\b")
    with open(sys.argv[1], "w") as f:
        f.write("SASS excerpts (cuobjdump -sass libquark_b200.so): every tcgen05 / TMEM / TMA instruction of the kernels that issue tcgen05.mma\n")
        cur, keep, lines = None, False, []
        for line in txt.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                keep = kern[cur]["UTCHMMA"] > 0
                if keep:
                    d = re.sub(r"\((int|bool|unsigned int)\)", "", demangle(cur))
                    f.write("\n== " + re.sub(r"\(.*", "", d) + f"  ({kern[cur]['_n']} instructions)\n")
                continue
            if keep and pat.search(line):
                f.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()) + "\n")
