"""Static check on the built library (no GPU): the shared-memory stage release of every ring-buffer kernel still carries
its data dependency in SASS.

`warp_release_after_loads` (csrc/common.cuh) predicates the mbarrier arrive on `REDUX.OR(dep) != never`, `never` being
a launch parameter.  ptxas folded an earlier form that compared against a literal (the arrive then no longer waited for
the ld.shared it releases; profiles/r02_mc_single_pixel.txt) -- nothing fails at run time when that happens, results
are just wrong once in a few hundred launches.  So the instruction pattern itself is asserted here:

    REDUX.OR URx, Ry ; ... ISETP.NE.U32 ... , UR<param> ... ; @P SYNCS.ARRIVE.TRANS64.A1T0 ...
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "semanticsegmentationactivelearning_b200", "libalscore.so")
RING_KERNELS = ("score_tiles_kernel", "mc_update_kernel", "score_head_kernel")


@pytest.fixture(scope="module")
def functions():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = out.setdefault(m.group(1), [])
        elif cur is not None and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            cur.append(line)
    return out


def test_every_ring_kernel_releases_its_stage_behind_the_loads(functions):
    ring = {k: v for k, v in functions.items() if any(t in k for t in RING_KERNELS)}
    assert len(ring) >= 400, "expected the scoring, streamed-update and fused-head instantiations, found %d" % len(ring)
    bad = []
    for name, lines in ring.items():
        redux = [i for i, l in enumerate(lines) if "REDUX.OR" in l]
        if len(redux) < 1:
            bad.append((name, "no REDUX.OR: the dependency word is dead code"))
            continue
        for r in redux:
            m = re.search(r"REDUX\.OR (UR\d+),", lines[r])
            arrive = next((i for i in range(r, len(lines)) if "SYNCS.ARRIVE.TRANS64.A1T0" in lines[i]), None)
            if arrive is None:
                bad.append((name, "no stage release behind the reduction"))
                continue
            between = lines[r:arrive + 1]
            # the reduced word reaches a compare against a uniform register (the `never` launch parameter) ...
            # (ptxas writes it as NE.AND or, merged with the lane test, as EQ.OR)
            cmp_ok = any(re.search(r"ISETP\.(NE|EQ)\.U32\.(AND|OR) P\d, PT, R\d+, UR\d+, ", l) for l in between)
            # ... and the arrive is predicated
            pred_ok = re.search(r"@!?P\d\s+SYNCS\.ARRIVE", lines[arrive]) is not None
            if not (m and cmp_ok and pred_ok):
                bad.append((name, "release not predicated on the reduced word: " + lines[arrive].strip()[:80]))
    assert not bad, "%d of %d kernels: %s" % (len(bad), len(ring), bad[:3])
