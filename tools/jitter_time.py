"""One-time cost of the MT19937 snapshot walk for a 4K x 4 spp frame (wall clock around the first jitter request)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
rt = g.load_rtds()
ctx = rt.Rtds(0)
ctx.jitter_stream(0, 16)
t0 = time.perf_counter(); ctx.jitter_stream(66_355_000, 16); t1 = time.perf_counter()
print("snapshot walk to double 66.4M: %.2f ms" % ((t1 - t0) * 1e3))
