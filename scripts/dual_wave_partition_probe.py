"""Dev tool: one rank's share of an N-GPU frame (tile partition rank 0 of N) rendered on one GPU with one / two waves in flight."""
import hashlib
import os
import subprocess
import sys

if len(sys.argv) > 2:
    sys.path.insert(0, ".")
    from foundation_b200 import pt, scenes  # noqa: E402
    sc = scenes.fractal_terrain()
    n = int(sys.argv[1])
    tr = pt.PathTracer(sc.width, sc.height, background=sc.background)
    tr.partition_set(0, n, 32)
    tr.load(sc)
    tr.render(0, 16, 8)
    best = 1e9
    for _ in range(3):
        tr.render(0, 64, 8); best = min(best, tr.stats().last_ms)
    print(f"rank 0 of {n}, {sys.argv[2]}: {best:.2f} ms for 64 spp of the owned tiles  frame {hashlib.sha256(tr.read_accum().tobytes()).hexdigest()[:12]}", flush=True)
else:
    for n in (8, 4):
        for dual in (0, 1):
            subprocess.run([sys.executable, __file__, str(n), f"dual={dual}"], env=dict(os.environ, FOUNDATION_PT_DUAL_WAVE=str(dual)), timeout=200)
