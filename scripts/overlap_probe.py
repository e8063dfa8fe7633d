"""Dev tool: upper bound of what overlapping two independent wavefront renders on one GPU buys (two contexts, two streams) against one context rendering all samples."""
import sys
import time

sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "terrain"
sc = scenes.by_name(name)
trs = [pt.PathTracer(sc.width, sc.height, background=sc.background) for _ in range(2)]
for t in trs:
    t.load(sc)
    t.render(0, 16, 8)
for rep in range(3):
    t0 = time.perf_counter(); trs[0].render(0, 64, 8); a = time.perf_counter() - t0
    t0 = time.perf_counter()
    trs[0].render_async(0, 32, 8); trs[1].render_async(32, 32, 8)
    trs[0].wait(); trs[1].wait()
    b = time.perf_counter() - t0
    print(f"{name}: one context 64 spp {a * 1e3:.1f} ms ({64 / a:.1f} spp/s); two contexts 32 + 32 spp concurrently {b * 1e3:.1f} ms ({64 / b:.1f} spp/s): {100 * (a / b - 1):+.1f} %", flush=True)
