"""One rank of tests/test_gpu_multi.py::test_one_process_per_gpu_comm_init_and_gather: argv = rank world mode id_file out_file.
The 128-byte communicator id travels through a file (any channel works: the C ABI only needs the bytes)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foundation_b200 import pt  # noqa: E402
from tests.util import SMALL_SCENES  # noqa: E402

rank, world, mode, idf, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
sc = SMALL_SCENES["terrain"]()
tr = pt.PathTracer(sc.width, sc.height, device=rank, seed=13, background=sc.background)
tr.load(sc)
if rank == 0:
    cid = pt.PathTracer.comm_unique_id()
    with open(idf + ".tmp", "wb") as f:
        f.write(cid)
    os.replace(idf + ".tmp", idf)
else:
    t0 = time.time()
    while not os.path.exists(idf):
        if time.time() - t0 > 300:
            raise SystemExit("no communicator id")
        time.sleep(0.05)
    cid = open(idf, "rb").read()
tr.comm_init(cid, rank, world, 16, pt.COMM_DIRECT if mode == "direct" else 0)
tr.render(0, 1, 4); tr.render(1, 2, 4)
tr.gather(0)
if rank == 0:
    np.save(out, tr.read_accum())
tr.gather(0)          # a second gather is idempotent and keeps every rank alive until the root has read its frame
tr.close()
