"""Host logic of the denoising loop: the folded DPM-Solver++ coefficients the fused kernel consumes reproduce the
oracle's (diffusers-restated) scheduler step on CPU tensors."""
import torch

from ecad_b200.pipeline import DPMSolverPP2M
from oracle.pixart_oracle import OracleDPMSolver


def test_coefficients_reproduce_oracle_solver():
    n = 20
    prod, ref = DPMSolverPP2M(), OracleDPMSolver(n)
    prod.set_timesteps(n)
    assert prod.timesteps.tolist() == ref.timesteps.tolist()
    assert torch.allclose(torch.from_numpy(prod.sigmas).float(), ref.sigmas, rtol=1e-6, atol=0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 8, 8, generator=g)
    x_ref = x.clone()
    x0_prev = torch.zeros_like(x)
    for i in range(n):
        eps = torch.randn(2, 4, 8, 8, generator=g)
        c = prod.coefficients()
        x0 = (x - c["sigma_s"] * eps) / c["alpha_s"]
        x = c["c_x"] * x + c["c_d0"] * x0 + c["c_d1"] * x0_prev
        x0_prev = x0
        prod.advance()
        x_ref = ref.step(eps, x_ref)
        # random eps makes the iterates grow to O(100); compare relative to the iterate's scale (fp32 round-off)
        assert float((x - x_ref).abs().max() / x_ref.abs().max()) < 2e-5, i
    # order pattern: first and last updates are first-order
    prod.set_timesteps(n)
    assert prod.coefficients()["c_d1"] == 0.0


def test_pipeline_argument_validation_without_gpu():
    import pytest

    from ecad_b200.pipeline import B200PixArtPipeline

    class FakeTr:
        device = torch.device("cpu")

    p = B200PixArtPipeline(FakeTr())
    with pytest.raises(ValueError, match="text encoding is out of scope"):
        p(prompt="a cat")
    with pytest.raises(ValueError, match="required"):
        p()
