"""SASS evidence: per-kernel histogram of the opcodes that show what the kernels are made of (B200_PROFILING.md,
"What proves a Blackwell-native kernel"), generated from the built library with cuobjdump.

    python profiles/sass_opcodes.py [path/to/libalscore.so] > profiles/sass_opcodes.txt

UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG = cp.async.bulk.tensor (TMA tile load),
UBLKCP = cp.async.bulk (1-D bulk copy, TMA engine), SYNCS = mbarrier, FFMA2 / FADD2 = packed fp32 pairs,
FMNMX3 = 3-input min/max, MUFU = ex2 / lg2 / rcp, REDG = global reduction / atomic (RED, REDG, ATOMG: per-image fixed-point sums, tile counter),
UTCBAR = tcgen05.commit, USETMAXREG = setmaxnreg (bring-up builds only)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "semanticsegmentationactivelearning_b200", "libalscore.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMNMX3", "MUFU", "REDG", "LDS",
        "STG", "HMMA", "total"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    op_re = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)")
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = op_re.match(line)
        if m:
            op = m.group(1)
            if op in ("RED", "REDG", "ATOMG", "ATOM", "REDUX"):
                op = "REDG"       # global reductions / atomics, whichever form ptxas picked
            cur[op] += 1
            cur["total"] += 1
    pretty = demangle(list(kernels))
    print("# SASS opcode histogram of %s (%d kernels), cuobjdump -sass" % (os.path.relpath(LIB, ROOT), len(kernels)))
    print("# " + __doc__.split("\n\n")[2].replace("\n", "\n# "))
    # families
    fam = collections.OrderedDict()
    for name, cnt in kernels.items():
        p = pretty[name]
        f = re.sub(r"<.*", "", p.replace("void ", "").replace("als::", ""))
        agg = fam.setdefault(f, [0, collections.Counter()])
        agg[0] += 1
        agg[1].update(cnt)
    print("\n## per kernel family (summed over its instantiations)")
    print("%-28s %5s " % ("family", "n") + " ".join("%8s" % k for k in KEYS))
    for f, (n, cnt) in fam.items():
        print("%-28s %5d " % (f[:28], n) + " ".join("%8d" % cnt.get(k, 0) for k in KEYS))
    print("\n## the instantiations the BASELINE workloads launch")
    want = [r"score_tiles_kernel<float, 19, 0>", r"score_tiles_kernel<float, 19, 1>", r"score_tiles_kernel<float, 19, 4>",
            r"score_tiles_kernel<float, 6, 0>", r"score_tiles_kernel<float, 66, 4>", r"score_tiles_kernel<__nv_bfloat16, 19, 0>",
            r"score_tiles_kernel<__nv_bfloat16, 6, 0>", r"score_head_kernel<19, 0, 4>", r"score_head_kernel<19, 1, 4>",
            r"score_head_kernel<19, 4, 2>", r"score_head_kernel<6, 0, 4>", r"mc_update_kernel<float, 19, true>",
            r"mc_update_kernel<float, 19, false>", r"mc_finish_kernel<float, 19>", r"select_kernel", r"rank_scatter_kernel",
            r"finalize_kernel", r"score_generic_kernel<float>", r"synth_kernel<float>"]
    print("%-52s " % "kernel" + " ".join("%8s" % k for k in KEYS))
    for w in want:
        for name, cnt in kernels.items():
            p = pretty[name].replace("(int)", "").replace("(bool)", "").replace("als::", "").replace("void ", "")
            if w in p:
                print("%-52s " % w[:52] + " ".join("%8d" % cnt.get(k, 0) for k in KEYS))
                break
        else:
            print("%-52s  (not found)" % w)


if __name__ == "__main__":
    main()
