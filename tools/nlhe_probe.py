"""NLHE MCCFR timing probe: python tools/nlhe_probe.py [batch ...] — per-phase device ms and updates/s (not a bench line)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robopoker_b200.nlhe import Nlhe  # noqa: E402

for batch in [int(x) for x in sys.argv[1:]] or [1024, 16384, 65536]:
    g = Nlhe(batch=batch, seed=1, table_slots=1 << 24)
    g.step(3)
    c0 = g.counters()
    ms = g.step_timed(5, flush_l2=True)
    c1 = g.counters()
    up = c1["updates"] - c0["updates"]
    print(batch, [round(x / 5, 3) for x in ms], "updates/epoch", up // 5, "updates/s %.3e" % (up / (ms[0] * 1e-3)), c1, flush=True)
