"""cProfile of the eager link-prediction step (FB15k-237 shape) to see where the host time of `e2e` goes."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
sys.argv = ["bench.py", "--shape", "fb15k237", "--no-cpu-baseline", "--steps", "3", "--warmup", "1"]
pr = cProfile.Profile()
orig_timed = None
pr.enable()
try:
    bench.main()
finally:
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue()[:9000])
