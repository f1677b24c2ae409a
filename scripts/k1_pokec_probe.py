"""Breakdown of the tensor-core stage at pokec scale: python scripts/k1_pokec_probe.py [nq]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.k1_probe import run
N = 1632803
nq = int(sys.argv[1]) if len(sys.argv) > 1 else N
mode = sys.argv[2] if len(sys.argv) > 2 else "all"
if mode in ("all", "fast"):
    run(N, 65, "clustered", 2.0, 4, 1, cand=14, reps=2, nq=nq)                                  # nothing passes: pure fast path
if mode in ("all", "seed"):
    run(N, 65, "clustered", -0.0011, 4, 1, cand=14, reps=2, nq=nq, seed_stride=16, seed_q=6)
if mode in ("all", "noseed"):
    run(N, 65, "clustered", -0.0011, 4, 1, cand=14, reps=2, nq=nq)
