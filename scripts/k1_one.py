"""One stage-1 launch (for ncu): python scripts/k1_one.py N d kind thr_lo mb nsplit [cand]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.k1_probe import run
a = sys.argv
run(int(a[1]), int(a[2]), a[3], float(a[4]), int(a[5]), int(a[6]), cand=int(a[7]) if len(a) > 7 else 32, reps=1)
