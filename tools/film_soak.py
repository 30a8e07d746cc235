"""Soak run of the film fuzz: python tools/film_soak.py FIRST COUNT — random resolution / crop / filter / radius / tiles / spp / wave /
rank configurations (tests/test_gpu_round2.py::test_random_film_configurations_match_the_oracle) against the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from arendur_b200 import api
import test_gpu_round2 as T
first, count = int(sys.argv[1]), int(sys.argv[2])
ctx = api.Context(0)
bad = []; t0 = time.time()
for seed in range(first, first + count):
    try:
        T.test_random_film_configurations_match_the_oracle(ctx, seed)
    except AssertionError as e:
        bad.append(seed); print(f"seed {seed}: {str(e)[:400]}", flush=True)
print(f"film soak: {count} configurations, {len(bad)} with differences {bad[:20]}, {time.time() - t0:.1f} s")
