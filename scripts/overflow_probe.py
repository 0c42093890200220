"""Which count path reports FKM_EOVERFLOW for a homopolymer of 2^32 + 5000 bases (test_count_overflow_is_reported, per mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fastkmer_b200 as fk
from fastkmer_b200 import api
n_pos = (1 << 32) + 5000
nw = (n_pos + 31) // 32
bases = np.zeros(nw, dtype=np.uint64); inv = np.zeros(nw, dtype=np.uint32)
c2 = fk.Context(0)
for ht, mode in ((1, 0), (1, 1), (1, 2), (0, 0)):
    c2.set("count_mode", mode)
    cfg = fk.TestConfiguration("", "", 28, 10, 3, max_b=2048, useHT=bool(ht), write=False)
    try:
        res, st = c2.count_packed_host(cfg, bases, inv, n_pos, want_result=False)
        print(ht, mode, "NO ERROR", st["n_kmers"], st["total_count"], st["n_distinct"], st["n_fallbacks"], st["n_mid_bins"], flush=True)
    except api.FkmError as e:
        print(ht, mode, "error", e.code, str(e)[:100], flush=True)
c2.close()
