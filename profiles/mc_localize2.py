"""Bring-up: WHICH pixels differ between repeated streamed runs (k samples, confidence map)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from semanticsegmentationactivelearning_b200 import Scorer
sc = Scorer(0)
shape = (2, 512, 1024, 19)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
x8 = sc.synth_logits(8, 3, *shape)
ref = sc.pseudo_annotation(x8[:k].contiguous(), "confidence")["pseudo_confidence"].clone()
for rep in range(reps):
    sc.mc_begin(shape)
    for t in range(k):
        sc.mc_add_sample(x8[t])
    m = sc.mc_finish("confidence", want_maps=True)["pseudo_confidence"]
    d = (m != ref).flatten().nonzero().flatten().cpu().numpy()
    if d.size:
        tiles = np.unique(d // 256)
        print("rep %d: %d pixels differ in %d tiles; tiles %s; offsets in tile: min %d max %d; first pixels %s; got %s want %s" % (
            rep, d.size, tiles.size, tiles[:8].tolist(), (d % 256).min(), (d % 256).max(), d[:6].tolist(),
            m.flatten()[d[:3]].tolist(), ref.flatten()[d[:3]].tolist()))
print("done")
