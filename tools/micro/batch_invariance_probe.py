"""Batch invariance of a cached config-2 generation (the check of tests/test_gpu_fullsize.py, without the CPU oracle):
100 prompts in one batch against the same prompts four at a time, and the batch-100 run repeated (determinism).
    [ECAD_B200_LIB=...] python tools/micro/batch_invariance_probe.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from ecad_b200.image_generator import B200PixArtAlphaImageGenerator  # noqa: E402
from ecad_b200.schedule import load_packed_schedules, schedule_from_packed  # noqa: E402
from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings  # noqa: E402

rows = load_packed_schedules(ROOT / "tests" / "golden" / "pixart_schedules.json.gz")
row = [r for r in rows if r["path"].endswith("gen_000/candidates/cand_003.json")][0]
sd = random_init_state_dict(PixArtConfig(), 0)
B = 100
emb = synthetic_prompt_embeddings(B, seed=1)
gen = B200PixArtAlphaImageGenerator(cache_schedule=schedule_from_packed(row), start_seed=0, state_dict=sd)
full = gen.generate_images(emb)[0].cpu()
again = gen.generate_images(emb)[0].cpu()
print("batch 100 repeat: max abs diff", float((full - again).abs().max()), "finite", bool(torch.isfinite(full).all()))
noise = torch.randn(B, 4, 32, 32, generator=torch.Generator().manual_seed(0))
steps = len(row["flags"]) if "flags" in row else 20
for lo in (0, 48, 96):
    part = {k: v[lo:lo + 4] for k, v in emb.items()}
    small = gen.diffusion_pipeline(
        prompt_embeds=part["prompt_embeds"], prompt_attention_mask=part["prompt_attention_mask"],
        negative_prompt_embeds=part["negative_prompt_embeds"],
        negative_prompt_attention_mask=part["negative_prompt_attention_mask"], latents=noise[lo:lo + 4].clone(),
        num_inference_steps=20, callback=gen._call_callbacks_wrapper)[0].cpu()
    rel = float((small - full[lo:lo + 4]).abs().max() / full.abs().max())
    per = [(float((small[i] - full[lo + i]).abs().max() / full.abs().max())) for i in range(4)]
    print(f"lo={lo}: rel {rel:.2e} per-prompt {['%.1e' % x for x in per]}")
