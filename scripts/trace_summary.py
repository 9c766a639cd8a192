"""Developer tool: summarise a BP4_TRACE file (library built with -DBP4_PHASE_TIMING, run with
BP4_TRACE=<file>): mean time between the stamps of a batch, over the blocks and batches of every
traced launch.

  python scripts/trace_summary.py FILE
"""
import sys

import numpy as np

SLOTS = 20
# stamps in program order and what ends at each of them
ORDER = [(0, "top (after the barrier that closes the previous iteration)"), (5, "request of pre job A + phase 1"),
         (6, "pre job A"), (7, "barrier + phase 2"), (8, "pre job B"), (9, "barrier + phase 3 + park"),
         (10, "post job A"), (1, "barrier"), (12, "gather issue"), (2, "scatter"), (3, "barrier"),
         (18, "gather store"), (13, "post job B")]

raw = np.fromfile(sys.argv[1], dtype=np.uint64)
pos, launch = 0, 0
while pos < raw.size:
    assert raw[pos] == 0xB4B4B4B4
    grid, nb, fused = int(raw[pos + 1]), int(raw[pos + 2]), int(raw[pos + 3])
    t = raw[pos + 4:pos + 4 + grid * nb * SLOTS].reshape(grid, nb, SLOTS).astype(np.int64)
    pos += 4 + grid * nb * SLOTS
    launch += 1
    ok = (t[:, :, 0] > 0) & (t[:, :, 3] > 0)
    ok[:, 0] = False  # the first batch carries the prologue
    if ok.sum() == 0:
        continue
    print(f"launch {launch}: grid {grid}, fused {fused}, {int(ok.sum())} traced batches")
    prev = None
    for k, name in ORDER:
        if prev is not None:
            good = ok & (t[:, :, k] > 0) & (t[:, :, prev] > 0)
            if good.sum():
                d = (t[:, :, k] - t[:, :, prev])[good] * 1e-3
                print(f"   {name:48s} {d.mean():7.2f} us  (p10 {np.percentile(d,10):6.2f}, p90 {np.percentile(d,90):6.2f})")
                prev = k
        else:
            prev = k
    for a, b, name in [(5, 16, "pre job A: wait"), (7, 17, "pre job B: wait"), (9, 14, "post job A: wait"),
                       (18, 15, "post job B: wait")]:
        good = ok & (t[:, :, a] > 0) & (t[:, :, b] > 0)
        if good.sum():
            d = (t[:, :, b] - t[:, :, a])[good] * 1e-3
            print(f"   {name:48s} {d.mean():7.2f} us  in {100.0*good.sum()/ok.sum():.0f} % of the batches")
    nxt = (t[:, 1:, 0] - t[:, :-1, 0])[ok[:, :-1] & (t[:, 1:, 0] > 0)] * 1e-3
    print(f"   {'period per batch':48s} {nxt.mean():7.2f} us")
