"""Runs the randomised GPU-vs-oracle comparison of test_gpu_fuzz.py over many more seeds and prints the worst
difference (test infrastructure: it imports the oracle, so it lives under tests/).  python tests/fuzz_stats.py"""
import sys, os, warnings
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import numpy as np
import svgrasterize_b200 as B
from oracle import render as O
from svgrasterize_b200 import scene as S, synth
import test_gpu_fuzz as F
warnings.simplefilter("ignore")
worst=0; n2=0; nz=[]; errs=0; typeerr=0
for seed in range(48, 400):
    rng = np.random.default_rng(7000 + seed)
    size = (int(rng.integers(48, 140)), int(rng.integers(48, 140)))
    scale = float(rng.uniform(0.8, 2.2))
    scene = F._rand_scene(rng, S, synth).transform(S.Transform().scale(scale).rotate(float(rng.uniform(-0.2, 0.2))))
    lin = bool(seed % 3 == 0)
    try:
        ref = O.render_canvas(scene, size, lin)
    except TypeError:
        typeerr += 1
        try:
            B.render_canvas(scene, size, lin); print("seed", seed, "oracle TypeError but GPU rendered")
        except TypeError: pass
        continue
    except Exception as e:
        print("seed", seed, "oracle exception", type(e).__name__, e); errs+=1; continue
    try:
        got = B.render_canvas(scene, size, lin)
    except Exception as e:
        print("seed", seed, "GPU exception", type(e).__name__, str(e)[:100]); errs+=1; continue
    if ref is None:
        if got.any(): print("seed", seed, "ref None, gpu nonzero")
        continue
    d = np.abs(got.astype(int) - ref.astype(int))
    nz.append(float((ref[...,3] > 0).mean()))
    if d.max() > 1:
        n2 += 1; print("seed", seed, "max", int(d.max()), "count>1", int((d>1).sum()), "size", size)
    worst = max(worst, int(d.max()))
print("worst", worst, "seeds with >1:", n2, "mean coverage", np.mean(nz), "typeerr", typeerr, "errs", errs)
