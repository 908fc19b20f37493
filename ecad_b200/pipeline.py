"""Denoising loop around the B200 transformer - same contract as the reference's pipelines.

The reference drives the transformer from diffusers' ``PixArtAlphaPipeline.__call__`` (faithful copy with marked
edits at /root/reference/ecad/pipelines/pass_through.py:186-404; TGATE variant ecad/pipelines/tgate.py:327-423):
CFG concat -> ``transformer(...)`` -> CFG combine -> learned-sigma drop -> ``scheduler.step`` ->
``callback(step, t, latents)``.  diffusers is not installable here, so the loop is restated with the same argument
names and the same callback protocol; the VAE decode after the loop is out of scope (latents are the output,
SURVEY.md section 8d).  The per-step tail (CFG + sigma drop + DPM-Solver++ update) is one fused kernel.
"""
from __future__ import annotations

import math
from typing import Any, Callable

import numpy as np
import torch

from . import _lib


class DPMSolverPP2M:
    """DPM-Solver++(2M) with PixArt's scheduler_config (dpmsolver++, order 2, midpoint, epsilon prediction, linear
    betas 1e-4..0.02 over 1000 steps, linspace spacing, lower_order_final, final sigma 0) - the configuration
    diffusers' DPMSolverMultistepScheduler runs under in the reference.  Coefficients are
    folded on the host in float64; the update itself is `ecadk_cfg_dpm_step`."""

    order = 1  # scheduler.order as the pipeline's callback arithmetic sees it
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02):
        # fp32 betas / cumprod like diffusers, so the sigma table matches the reference's bit for bit
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self._train_sigmas = (((1 - alphas_cumprod) / alphas_cumprod) ** 0.5).numpy().astype(np.float64)
        self.num_train_timesteps = num_train_timesteps
        self.timesteps: torch.Tensor = torch.empty(0, dtype=torch.int64)
        self.sigmas = np.zeros(0)
        self.step_index = 0
        self.lower_order_nums = 0

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        ts = (np.linspace(0, self.num_train_timesteps - 1, num_inference_steps + 1).round()[::-1][:-1]
              .copy().astype(np.int64))
        sig = np.interp(ts, np.arange(0, len(self._train_sigmas)), self._train_sigmas)
        self.sigmas = np.concatenate([sig, [0.0]])
        self.timesteps = torch.from_numpy(ts)
        self.step_index = 0
        self.lower_order_nums = 0

    def scale_model_input(self, sample, timestep=None):
        return sample

    @staticmethod
    def _alpha_sigma(sigma: float) -> tuple[float, float]:
        alpha_t = 1.0 / math.sqrt(sigma * sigma + 1.0)
        return alpha_t, sigma * alpha_t

    def coefficients(self) -> dict[str, float]:
        """Folded coefficients of the update at ``step_index``: x_next = c_x*x + c_d0*x0 + c_d1*x0_prev."""
        i, n = self.step_index, len(self.timesteps)
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[i])
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i + 1])
        lam_s0 = math.log(alpha_s0) - math.log(sigma_s0)
        first_order = self.lower_order_nums < 1 or i == n - 1  # lower_order_final / final sigma zero
        if sigma_t == 0.0:
            a = -alpha_t  # exp(-h) -> 0 as lambda_t -> inf
            c_x = 0.0
        else:
            lam_t = math.log(alpha_t) - math.log(sigma_t)
            h = lam_t - lam_s0
            a = alpha_t * (math.exp(-h) - 1.0)
            c_x = sigma_t / sigma_s0
        if first_order:
            c_d0, c_d1 = -a, 0.0
        else:
            alpha_s1, sigma_s1 = self._alpha_sigma(self.sigmas[i - 1])
            lam_s1 = math.log(alpha_s1) - math.log(sigma_s1)
            r0 = (lam_s0 - lam_s1) / h
            c_d0, c_d1 = -a - 0.5 * a / r0, 0.5 * a / r0
        return {"sigma_s": sigma_s0, "alpha_s": alpha_s0, "c_x": c_x, "c_d0": c_d0, "c_d1": c_d1}

    def advance(self) -> None:
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self.step_index += 1


class B200PixArtPipeline:
    """``pipeline(prompt_embeds=..., negative_prompt_embeds=..., ..., callback=...)`` -> latents.

    Keyword names follow the reference's call site (ecad/image_generators/pixart_image_generator.py:358-383).
    ``gate_step`` enables the TGATE loop variant (ecad/pipelines/tgate.py:329-341,382-389).
    """

    def __init__(self, transformer, scheduler: DPMSolverPP2M | None = None, gate_step: int | None = None,
                 use_cuda_graph: bool = False):
        from .graphs import GenerationGraphs

        self.use_cuda_graph = use_cuda_graph
        self._graphs = GenerationGraphs()
        self._ts_cache = None
        self.transformer = transformer
        self.scheduler = scheduler if scheduler is not None else DPMSolverPP2M()
        self.gate_step = gate_step
        self.vae_scale_factor = 8
        self.device = transformer.device

    @classmethod
    def from_pretrained(cls, transformer, **kwargs) -> "B200PixArtPipeline":
        """The constructor under the name the reference's registry calls (load_pipeline.py:55-56): the "checkpoint" of
        this pipeline is the resident transformer - text encoder and tokenizer are out of scope, the VAE decoder
        belongs to the image generator."""
        return cls(transformer, **kwargs)

    def prepare_latents(self, batch_size, num_channels, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            # diffusers randn_tensor: a CPU generator draws on the CPU, then the tensor moves to the device
            gen_dev = generator.device if generator is not None else torch.device("cpu")
            latents = torch.randn(shape, generator=generator, device=gen_dev, dtype=torch.float32)
        latents = latents.to(device=device, dtype=torch.float32)
        return latents * self.scheduler.init_noise_sigma

    def _denoise(self, inp: dict[str, torch.Tensor], batch_size: int, do_cfg: bool, guidance_scale: float, height: int,
                 width: int, callback, callback_steps: int) -> torch.Tensor:
        """The denoising loop proper (pass_through.py:296-380); every tensor it touches is already on the device, so
        the whole loop can be recorded into one CUDA graph."""
        tr, sched, dev = self.transformer, self.scheduler, self.device
        cfgm = tr.config
        latents, embeds, mask = inp["latents"], inp["embeds"], inp["mask"]
        negative_prompt_embeds = inp["negative_prompt_embeds"]
        negative_prompt_attention_mask = inp["negative_prompt_attention_mask"]
        sched.step_index, sched.lower_order_nums = 0, 0
        latent_channels = cfgm.in_channels
        x0_prev = torch.zeros_like(latents)
        added_cond_kwargs = {"resolution": inp.get("resolution"), "aspect_ratio": inp.get("aspect_ratio")}
        lib = _lib.load()
        hw = latents.shape[-2] * latents.shape[-1]
        timesteps = sched.timesteps
        timesteps_dev = self._timesteps_on_device(timesteps)
        for i, t in enumerate(timesteps):
            gated = self.gate_step is not None and i >= self.gate_step
            cond_in = added_cond_kwargs
            if gated:
                model_in, e_in, m_in = latents, negative_prompt_embeds, negative_prompt_attention_mask
                cond_in = {k: (v[:batch_size] if v is not None else None) for k, v in added_cond_kwargs.items()}
            else:
                model_in = torch.cat([latents] * 2) if do_cfg else latents
                e_in, m_in = embeds, mask
            model_in = sched.scale_model_input(model_in, t)
            current_timestep = timesteps_dev[i:i + 1].expand(model_in.shape[0])
            if hasattr(tr, "hint_timestep"):
                tr.hint_timestep(float(t))  # host value of the shared timestep (table cache key), not an argument
            with _lib.nvtx_range(f"step {i:02d} transformer"):
                noise_pred = tr(model_in, encoder_hidden_states=e_in, encoder_attention_mask=m_in,
                                timestep=current_timestep, added_cond_kwargs=cond_in, return_dict=False)[0]
            c = sched.coefficients()
            _lib.check(
                lib.ecadk_cfg_dpm_step(noise_pred.data_ptr(), latents.data_ptr(), x0_prev.data_ptr(), batch_size,
                                       latent_channels, hw, int(do_cfg and not gated), float(guidance_scale),
                                       c["sigma_s"], c["alpha_s"], c["c_x"], c["c_d0"], c["c_d1"],
                                       _lib.stream_ptr()),
                "cfg_dpm_step")
            tr.launches += 1
            sched.advance()
            if callback is not None and i % callback_steps == 0:
                callback(i // getattr(sched, "order", 1), t, latents)
        return latents

    def _timesteps_on_device(self, timesteps: torch.Tensor) -> torch.Tensor:
        key = tuple(int(t) for t in timesteps)
        if self._ts_cache is None or self._ts_cache[0] != key:
            self._ts_cache = (key, timesteps.to(self.device))
        return self._ts_cache[1]

    @torch.no_grad()
    def __call__(
        self,
        prompt=None,
        negative_prompt=None,
        prompt_embeds: torch.Tensor | None = None,
        prompt_attention_mask: torch.Tensor | None = None,
        negative_prompt_embeds: torch.Tensor | None = None,
        negative_prompt_attention_mask: torch.Tensor | None = None,
        num_images_per_prompt: int = 1,
        num_inference_steps: int = 20,
        generator: torch.Generator | None = None,
        latents: torch.Tensor | None = None,
        guidance_scale: float = 4.5,
        height: int | None = None,
        width: int | None = None,
        callback: Callable[[int, Any, torch.Tensor], None] | None = None,
        callback_steps: int = 1,
        output_type: str = "latent",
        return_dict: bool = False,
        capture_callback: Callable[[int, Any, torch.Tensor], None] | None = None,
        **kwargs: Any,
    ):
        if prompt is not None or negative_prompt is not None:
            raise ValueError("text encoding is out of scope: pass prompt_embeds / negative_prompt_embeds")
        if prompt_embeds is None or negative_prompt_embeds is None:
            raise ValueError("prompt_embeds and negative_prompt_embeds are required")
        if output_type != "latent":
            raise NotImplementedError("VAE decode is out of scope; use output_type='latent'")
        if num_images_per_prompt != 1:
            raise NotImplementedError("the reference always calls with num_images_per_prompt=1")
        tr = self.transformer
        cfgm = tr.config
        dev = self.device
        height = height or cfgm.sample_size * self.vae_scale_factor
        width = width or cfgm.sample_size * self.vae_scale_factor
        batch_size = prompt_embeds.shape[0]
        do_cfg = guidance_scale > 1.0
        prompt_embeds = prompt_embeds.to(dev)
        negative_prompt_embeds = negative_prompt_embeds.to(dev)
        prompt_attention_mask = prompt_attention_mask.to(dev)
        negative_prompt_attention_mask = negative_prompt_attention_mask.to(dev)
        if do_cfg:
            embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
            mask = torch.cat([negative_prompt_attention_mask, prompt_attention_mask], dim=0)
        else:
            embeds, mask = prompt_embeds, prompt_attention_mask

        sched = self.scheduler
        sched.set_timesteps(num_inference_steps, device=dev)
        latent_channels = cfgm.in_channels
        latents = self.prepare_latents(batch_size, latent_channels, height, width, torch.float32, dev, generator,
                                       latents).contiguous()
        if cfgm.out_channels // 2 != latent_channels:
            raise NotImplementedError("only learned-sigma PixArt heads (out_channels = 2*in_channels) are supported")
        inputs = {"latents": latents, "embeds": embeds, "mask": mask, "negative_prompt_embeds": negative_prompt_embeds,
                  "negative_prompt_attention_mask": negative_prompt_attention_mask, "resolution": None,
                  "aspect_ratio": None}
        if cfgm.sample_size == 128 and getattr(tr.cfg, "resolved_additional_conditions", False):
            # 6.1 micro-conditions of the 1024-MS checkpoints (pass_through.py:268-290)
            resolution = torch.tensor([float(height), float(width)]).repeat(batch_size, 1)
            aspect_ratio = torch.tensor([float(height / width)]).repeat(batch_size, 1)
            if do_cfg:
                resolution = torch.cat([resolution, resolution], dim=0)
                aspect_ratio = torch.cat([aspect_ratio, aspect_ratio], dim=0)
            inputs["resolution"], inputs["aspect_ratio"] = resolution.to(dev), aspect_ratio.to(dev)
        if self.use_cuda_graph:
            # keyed on the schedule's CONTENT (not its id/name: in a set_schedule() loop over GA candidates a freed
            # schedule's address is reused) and dropped whenever the transformer's buffers moved (a recorded graph
            # holds raw pointers into the workspace and the per-timestep tables)
            if getattr(tr, "buffer_epoch", 0) != getattr(self, "_graph_epoch", None):
                self._graphs.clear()
                self._graph_epoch = getattr(tr, "buffer_epoch", 0)
            key = ("pixart", tr.cache_schedule.content_key(), tuple(latents.shape), tuple(embeds.shape),
                   num_inference_steps, float(guidance_scale), self.gate_step, height, width,
                   bool(getattr(tr, "skip_dead_cache_stores", True)))
            latents = self._graphs.run(
                key, inputs, lambda st, cb: self._denoise(st, batch_size, do_cfg, guidance_scale, height, width, cb, 1),
                capture_callback if capture_callback is not None else callback, tr)
            if getattr(tr, "buffer_epoch", 0) != self._graph_epoch:
                # the run's own eager warm-up re-allocated the workspace / per-timestep tables: the graph just recorded
                # points into the new buffers, every OLDER graph into freed memory
                self._graphs.keep_only(key)
                self._graph_epoch = getattr(tr, "buffer_epoch", 0)
            if callback is not None:  # user-visible per-step protocol (counters, extra callbacks, reset LAST)
                for i, t in enumerate(sched.timesteps):
                    if i % callback_steps == 0:
                        callback(i // getattr(sched, "order", 1), t, latents)
        else:
            latents = self._denoise(inputs, batch_size, do_cfg, guidance_scale, height, width, callback, callback_steps)
        if not return_dict:
            return (latents,)
        return {"images": latents}


class B200TGATEPipeline(B200PixArtPipeline):
    """The pipeline a schedule selects with ``config.pipeline = {"name": "tgate", "kwargs": {"gate_step": k}}``
    (/root/reference/ecad/pipelines/tgate.py:29-72): the same loop, with the half-batch variant from step ``k`` on.
    Like the reference's ``_post_init`` it refuses to be built without a gate step."""

    def __init__(self, transformer, scheduler: DPMSolverPP2M | None = None, gate_step: int | None = None,
                 use_cuda_graph: bool = False):
        if gate_step is None:
            raise ValueError("gate_step must be provided")  # tgate.py:51-55
        super().__init__(transformer, scheduler, gate_step=int(gate_step), use_cuda_graph=use_cuda_graph)

    @classmethod
    def from_pretrained(cls, transformer, gate_step: int | None = None, **kwargs) -> "B200TGATEPipeline":
        return cls(transformer, gate_step=gate_step, **kwargs)
