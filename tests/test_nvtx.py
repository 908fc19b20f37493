"""NVTX ranges ("NVTX ranges per sub-block" is a deliverable of the new build; the reference has
no tracing).  CPU part: the switch and the host-side context manager; GPU part: one range per executor call and per
EXECUTED sub-block, none for a reused one, and results that do not depend on the switch."""
import numpy as np
import pytest
import torch


def test_switch_and_noop_range_on_cpu():
    from ecad_b200 import _lib
    from ecad_b200.build import build_library

    build_library()
    before = _lib.nvtx_ranges()
    assert before >= 0
    was = _lib.nvtx_enabled()
    try:
        _lib.set_nvtx(False)
        with _lib.nvtx_range("off") as r:  # no tool attached, tracing off: nothing is pushed
            assert not r.live
        assert _lib.nvtx_ranges() == before  # only executor calls push library ranges
    finally:
        _lib.set_nvtx(was)


@pytest.mark.gpu
def test_one_range_per_executed_sub_block(cuda_device):
    from ecad_b200 import _lib
    from ecad_b200.schedule import PixArtCacheSchedule
    from ecad_b200.transformer import B200PixArtTransformer2D, SequentialDiTScheduler
    from ecad_b200.weights import PixArtConfig, random_init_state_dict, synthetic_prompt_embeddings

    cfg = PixArtConfig(num_layers=3)
    sd = random_init_state_dict(cfg, seed=0)
    flags = np.ones((2, 3, 3), bool)
    flags[1, 0, 0] = flags[1, 1, 2] = flags[1, 2, 1] = False  # step 1 reuses attn1 / ff / attn2 of blocks 0 / 1 / 2
    emb = synthetic_prompt_embeddings(1, seed=3)
    lat = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(5))
    x_in = torch.cat([lat, lat]).cuda()
    e_in = torch.cat([emb["negative_prompt_embeds"], emb["prompt_embeds"]]).cuda()
    m_in = torch.cat([emb["negative_prompt_attention_mask"], emb["prompt_attention_mask"]]).cuda()
    ts = torch.full((2,), 949, dtype=torch.int64).cuda()

    def two_steps():
        sched = PixArtCacheSchedule.from_numpy(flags, 2, 3)
        tr = B200PixArtTransformer2D(sd, cfg, SequentialDiTScheduler(2), sched)
        outs = []
        for step in range(2):
            outs.append(tr(x_in, encoder_hidden_states=e_in, encoder_attention_mask=m_in, timestep=ts,
                           added_cond_kwargs={"resolution": None, "aspect_ratio": None}, return_dict=False)[0].clone())
            sched.per_step_callback(step)
        return outs

    was = _lib.nvtx_enabled()
    try:
        _lib.set_nvtx(False)
        n0 = _lib.nvtx_ranges()
        plain = two_steps()
        assert _lib.nvtx_ranges() == n0
        _lib.set_nvtx(True)
        traced = two_steps()
        # per forward: 1 executor range + one per executed sub-block (9 at step 0, 6 at step 1)
        assert _lib.nvtx_ranges() - n0 == (1 + 9) + (1 + 6)
        with _lib.nvtx_range("host level") as r:
            assert r.live
    finally:
        _lib.set_nvtx(was)
    for a, b in zip(plain, traced):
        assert torch.equal(a, b)
