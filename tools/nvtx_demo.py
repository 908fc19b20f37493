"""Two forwards of a 3-block PixArt model with NVTX ranges on - the target of
`ncu --nvtx --nvtx-include "b01.ff]" ...` (tools/gpu_round2_c.sh): shows that a capture can be cut at the reference's
sub-block granularity.  Step 1 reuses block 0's attn1, block 1's ff and block 2's attn2 (no range, no launch)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from ecad_b200 import _lib  # noqa: E402
from ecad_b200.schedule import PixArtCacheSchedule  # noqa: E402
from ecad_b200.transformer import B200PixArtTransformer2D, SequentialDiTScheduler  # noqa: E402
from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings  # noqa: E402


def main() -> None:
    _lib.set_nvtx(True)
    cfg = PixArtConfig(num_layers=3)
    flags = np.ones((2, 3, 3), bool)
    flags[1, 0, 0] = flags[1, 1, 2] = flags[1, 2, 1] = False
    sched = PixArtCacheSchedule.from_numpy(flags, 2, 3)
    tr = B200PixArtTransformer2D(random_init_state_dict(cfg, 0), cfg, SequentialDiTScheduler(2), sched)
    emb = synthetic_prompt_embeddings(4, seed=3)
    lat = torch.randn(4, 4, 32, 32, generator=torch.Generator().manual_seed(5))
    x = torch.cat([lat, lat]).cuda()
    e = torch.cat([emb["negative_prompt_embeds"], emb["prompt_embeds"]]).cuda()
    m = torch.cat([emb["negative_prompt_attention_mask"], emb["prompt_attention_mask"]]).cuda()
    ts = torch.full((8,), 949, dtype=torch.int64).cuda()
    for step in range(2):
        with _lib.nvtx_range(f"step {step:02d} transformer"):
            tr(x, encoder_hidden_states=e, encoder_attention_mask=m, timestep=ts,
               added_cond_kwargs={"resolution": None, "aspect_ratio": None}, return_dict=False)
        sched.per_step_callback(step)
    torch.cuda.synchronize()
    print("library ranges pushed:", _lib.nvtx_ranges())


if __name__ == "__main__":
    main()
