"""Dev tool: upper bound of what overlapping independent wavefront renders on one GPU buys (k contexts, k streams) against one context rendering all samples."""
import sys
import time

sys.path.insert(0, ".")
from foundation_b200 import pt, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "terrain"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = scenes.by_name(name)
trs = [pt.PathTracer(sc.width, sc.height, background=sc.background) for _ in range(K)]
for t in trs:
    t.load(sc)
    t.render(0, 16, 8)
total = 96
for rep in range(3):
    t0 = time.perf_counter(); trs[0].render(0, total, 8); a = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i, t in enumerate(trs):
        t.render_async(i * (total // K), total // K, 8)
    for t in trs:
        t.wait()
    b = time.perf_counter() - t0
    print(f"{name}: one context {total} spp {a * 1e3:.1f} ms ({total / a:.1f} spp/s); {K} contexts x {total // K} spp concurrently {b * 1e3:.1f} ms ({total / b:.1f} spp/s): {100 * (a / b - 1):+.1f} %", flush=True)
