#!/usr/bin/env python
"""static SASS instruction mix per kernel of libsuma_b200.so (cuobjdump -sass): how many instructions of which class a
kernel consists of. Static counts (not executed counts), but they show what a kernel is made of -- e.g. the IEEE-exact
fp32 divides (MUFU.RCP + FFMA fix-up chains + slow-path calls) and polynomial transcendentals of the bit-exact contract.
usage: python profiles/sass_mix.py [lib.so] > profiles/rNN_sass_mix.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "semantic_suma_b200/lib/libsuma_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
CLASSES = [("fp32 FMA/MUL/ADD", r"^(FFMA|FMUL|FADD)\b"), ("fp32 other (FSETP/FSEL/FMNMX/FCHK...)", r"^F(SETP|SEL|MNMX|CHK|SET|RND)\b"),
           ("MUFU (rcp/rsq/sqrt...)", r"^MUFU"), ("fp64", r"^D(FMA|MUL|ADD|SETP)\b"), ("int IMAD/IADD3/LEA/LOP3/SHF", r"^(IMAD|IADD3|LEA|LOP3|SHF|IABS|IMNMX|ISETP|SEL|PRMT|POPC|FLO|BREV)\b"),
           ("conversions", r"^(I2F|F2I|F2F|I2I|I2FP|F2FP)\b"), ("global load", r"^(LDG|LD)\b"), ("global store", r"^(STG|ST)\b"),
           ("global atomic / reduction", r"^(ATOMG|ATOM|RED|REDG)\b"), ("shared load/store", r"^(LDS|STS|LDSM)\b"), ("shared atomic", r"^ATOMS\b"),
           ("local (stack) load/store", r"^(LDL|STL)\b"), ("warp shuffle / vote / match", r"^(SHFL|VOTE|VOTEU|MATCH|REDUX)\b"),
           ("barrier / sync", r"^(BAR|BSYNC|BSSY|WARPSYNC|MEMBAR|ERRBAR|DEPBAR)\b"), ("branch / call / exit", r"^(BRA|BRX|CALL|RET|EXIT|JMP|BREAK|BMOV)\b"),
           ("TMA / mbarrier", r"^(UTMALDG|UTMASTG|SYNCS|UBLKCP)"), ("uniform datapath", r"^U[A-Z]"), ("move / misc", r"^(MOV|S2R|S2UR|CS2R|NOP|R2UR|LDC|ULDC|PLOP3|P2R|R2P)\b")]
kern, mix, order = None, {}, []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1); mix[kern] = collections.Counter(); order.append(kern); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if kern and m:
        op = m.group(1)
        for name, pat in CLASSES:
            if re.match(pat, op):
                mix[kern][name] += 1; break
        else:
            mix[kern]["other"] += 1
        mix[kern]["_total"] += 1
print("# static SASS instruction mix, %s (cuobjdump -sass, sm_100a)" % lib)
for k in order:
    c = mix[k]; tot = c.pop("_total", 0)
    if tot < 200: continue
    short = re.sub(r"^_ZN2sb\d+", "", k)[:60]
    print("\n%s   (%d instructions)" % (short, tot))
    for name, n in c.most_common():
        print("  %5d  %4.1f %%  %s" % (n, 100.0 * n / tot, name))
