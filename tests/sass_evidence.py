"""Count the Blackwell-native SASS mnemonics per kernel in the built objects (cuobjdump -sass; no GPU needed):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, HMMA = legacy mma.sync (must be 0).
usage: python tests/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "pacingpseudo_b200", "_obj")
PAT = {"UTC*MMA (tcgen05.mma)": r"\bUTC\w*MMA", "LDTM (tcgen05.ld)": r"\bLDTM", "STTM (tcgen05.st)": r"\bSTTM",
       "UTMALDG (TMA load)": r"\bUTMALDG", "UTMASTG (TMA store)": r"\bUTMASTG", "UBLKCP": r"\bUBLKCP",
       "SYNCS (mbarrier)": r"\bSYNCS", "HMMA (legacy mma.sync)": r"\bHMMA", "MUFU.EX2": r"MUFU\.EX2",
       "MUFU.LG2": r"MUFU\.LG2", "MUFU.RCP": r"MUFU\.RCP"}


def main():
    print("# cuobjdump -sass of pacingpseudo_b200/_obj/*.o (sm_100a): Blackwell-native instruction counts per kernel")
    print("# kernels without any of these instructions are omitted; HMMA (legacy tensor path) must not appear\n")
    for obj in ("conv_tc.o", "conv_halo.o", "loss.o", "ops.o"):
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
        name, counts, total = None, collections.OrderedDict(), collections.Counter()
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"\(.*", "", name).replace("void ", "").replace("pp::", "")
                counts[name] = collections.Counter()
                continue
            if name is None or not re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
                continue
            total[name] += 1
            for key, pat in PAT.items():
                if re.search(pat, line):
                    counts[name][key] += 1
        print("## %s" % obj)
        for k, c in counts.items():
            keys = [x for x in c if not x.startswith("MUFU")]
            show = c if (keys or "lean" in k or "scribble_loss" in k) else None
            if show:
                print("%-62s %5d instr  %s" % (k[:62], total[k], ", ".join("%s %d" % (a.split(" ")[0], b) for a, b in c.items())))
        print()


if __name__ == "__main__":
    main()
