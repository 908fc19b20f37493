"""How the CPU oracle's images/s depends on the prompts per step (the bounded sample of `bench.py --impl reference`):
    python tools/cpu_batch_sensitivity.py [prompt counts ...]      (default 1 2 4)
One warm-up image, then one timed 20-step generation per prompt count under the paper's ours_fast schedule."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

head_row, _, _ = bench.load_candidates()
counts = [int(a) for a in sys.argv[1:]] or [1, 2, 4]
for r in counts:
    t0 = time.perf_counter()
    times, cores = bench.cpu_oracle_images_per_s([head_row, head_row], warmup=1, prompts=r)
    print(f"prompts/step {r:3d}: {r / times[0]:.4f} images/s  ({times[0]:.1f} s per step, {cores} threads, "
          f"wall {time.perf_counter() - t0:.0f} s)", flush=True)
