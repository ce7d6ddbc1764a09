"""Bring-up: after how many streamed samples does run-to-run variation appear?  (finish after k samples, `confidence`)"""
import sys
import torch
sys.path.insert(0, ".")
from semanticsegmentationactivelearning_b200 import Scorer
sc = Scorer(0)
shape = (2, 512, 1024, 19) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split("x"))
x8 = sc.synth_logits(8, 3, shape[0], shape[1], shape[2], shape[3])
for k in (1, 2, 3, 8):
    for measure in ("confidence", "entropy"):
        seen = {}
        for _ in range(30):
            sc.mc_begin(shape)
            for t in range(k):
                sc.mc_add_sample(x8[t])
            v = sc.mc_finish(measure).cpu().numpy().tobytes()
            seen[v] = seen.get(v, 0) + 1
        ref = sc.score(x8[:k].contiguous(), measure) if k > 1 else sc.score(x8[0], measure)
        sc.mc_begin(shape)
        for t in range(k):
            sc.mc_add_sample(x8[t])
        eq = bool(torch.equal(sc.mc_finish(measure), ref))
        print("k=%d %-10s distinct=%s  == resident: %s" % (k, measure, sorted(seen.values(), reverse=True), eq))
