"""FLUX denoising loop around B200FluxTransformer2D - the contract of diffusers' ``FluxPipeline.__call__`` as the
reference drives it (/root/reference/ecad/image_generators/flux_image_generator.py:330-352): packed 2x2 latents,
position ids, FlowMatchEulerDiscreteScheduler with FLUX's resolution-dependent time shift, distilled guidance (no CFG
pair), ``callback_on_step_end(pipeline, step, timestep, callback_kwargs)``.  Text encoders and the VAE are out of scope:
the inputs are prompt embeddings, the output is the packed latents (``output_type="latent"``).
"""
from __future__ import annotations

import math
from typing import Any, Callable

import numpy as np
import torch

from . import _lib


def calculate_shift(image_seq_len: int, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.16) -> float:
    """diffusers pipeline_flux.calculate_shift."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    return image_seq_len * m + (base_shift - m * base_seq_len)


class FlowMatchEulerDiscrete:
    """FlowMatchEulerDiscreteScheduler under FLUX.1-dev's scheduler_config (use_dynamic_shifting: sigma' =
    e^mu / (e^mu + (1/sigma - 1))); ``step`` is x += (sigma_next - sigma) * v, run by ``ecadk_axpy_f32``.

    ``config`` carries the values diffusers' FluxPipeline hands to ``calculate_shift`` (``scheduler.config.*``):
    FLUX.1-dev's scheduler_config.json = base_shift 0.5, max_shift 1.15, base_image_seq_len 256, max_image_seq_len 4096
    (also the constructor defaults of FlowMatchEulerDiscreteScheduler; the 1.16 in ``calculate_shift``'s signature is
    never used by the pipeline).  Recalled, not verifiable here: the reference ships no scheduler config."""

    order = 1
    num_train_timesteps = 1000

    def __init__(self, base_shift: float = 0.5, max_shift: float = 1.15, base_image_seq_len: int = 256,
                 max_image_seq_len: int = 4096):
        from types import SimpleNamespace

        self.config = SimpleNamespace(base_shift=base_shift, max_shift=max_shift,
                                      base_image_seq_len=base_image_seq_len, max_image_seq_len=max_image_seq_len,
                                      num_train_timesteps=self.num_train_timesteps, use_dynamic_shifting=True)
        self.sigmas = np.zeros(0, dtype=np.float32)
        self.timesteps = torch.empty(0)
        self.step_index = 0

    def set_timesteps(self, num_inference_steps: int | None = None, device=None, sigmas=None, mu: float | None = None):
        if sigmas is None:
            sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
        sigmas = np.asarray(sigmas, dtype=np.float64)
        if mu is None:
            raise ValueError("you have to pass a value for `mu` when `use_dynamic_shifting` is set to be `True`")
        sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1))
        sig32 = sigmas.astype(np.float32)
        self.timesteps = torch.from_numpy(sig32 * self.num_train_timesteps)
        self.sigmas = np.concatenate([sig32, np.zeros(1, dtype=np.float32)])
        self.step_index = 0

    def step_coefficient(self) -> float:
        return float(self.sigmas[self.step_index + 1]) - float(self.sigmas[self.step_index])

    def advance(self) -> None:
        self.step_index += 1


def pack_latents(latents: torch.Tensor) -> torch.Tensor:
    """FluxPipeline._pack_latents: [B, C, H, W] -> [B, (H/2)*(W/2), C*4]."""
    b, c, h, w = latents.shape
    latents = latents.view(b, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5)
    return latents.reshape(b, (h // 2) * (w // 2), c * 4)


def latent_image_ids(batch: int, h2: int, w2: int) -> torch.Tensor:
    """FluxPipeline._prepare_latent_image_ids (diffusers 0.30.3: repeated over the batch): [B, h2*w2, 3]."""
    ids = torch.zeros(h2, w2, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(h2)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w2)[None, :]
    return ids.reshape(1, h2 * w2, 3).repeat(batch, 1, 1)


class B200FluxPipeline:
    vae_scale_factor = 16  # FluxPipeline: 2 ** len(vae.config.block_out_channels) for the FLUX VAE

    def __init__(self, transformer, scheduler: FlowMatchEulerDiscrete | None = None, use_cuda_graph: bool = False):
        from .graphs import GenerationGraphs

        self.use_cuda_graph = use_cuda_graph
        self._graphs = GenerationGraphs()
        self.transformer = transformer
        self.scheduler = scheduler if scheduler is not None else FlowMatchEulerDiscrete()
        self.device = transformer.device

    @classmethod
    def from_pretrained(cls, transformer, **kwargs) -> "B200FluxPipeline":
        """See B200PixArtPipeline.from_pretrained (load_pipeline.py:55-56)."""
        return cls(transformer, **kwargs)

    def prepare_latents(self, batch_size, num_channels_latents, height, width, generator, latents=None):
        h = 2 * (int(height) // self.vae_scale_factor)
        w = 2 * (int(width) // self.vae_scale_factor)
        ids = latent_image_ids(batch_size, h // 2, w // 2)
        if latents is not None:
            return latents.to(device=self.device, dtype=torch.float32), ids
        gen_dev = generator.device if generator is not None else torch.device("cpu")
        latents = torch.randn((batch_size, num_channels_latents, h, w), generator=generator, device=gen_dev,
                              dtype=torch.float32)
        return pack_latents(latents).to(self.device).contiguous(), ids

    def _denoise(self, inp: dict[str, torch.Tensor], text_ids, img_ids, callback_on_step_end) -> torch.Tensor:
        """The loop of FluxPipeline.__call__; all tensors already on the device (recordable into one CUDA graph)."""
        tr, sched = self.transformer, self.scheduler
        latents, prompt_embeds = inp["latents"], inp["prompt_embeds"]
        batch_size = latents.shape[0]
        lib = _lib.load()
        sched.step_index = 0
        for i, t in enumerate(sched.timesteps):
            timestep = inp["timesteps"][i:i + 1].expand(batch_size)
            with _lib.nvtx_range(f"step {i:02d} transformer"):
                noise_pred = tr(hidden_states=latents, timestep=timestep, guidance=inp["guidance"],
                                pooled_projections=inp["pooled_prompt_embeds"], encoder_hidden_states=prompt_embeds,
                                txt_ids=text_ids, img_ids=img_ids, joint_attention_kwargs=None, return_dict=False)[0]
            _lib.check(lib.ecadk_axpy_f32(latents.data_ptr(), noise_pred.data_ptr(), sched.step_coefficient(),
                                          latents.numel(), _lib.stream_ptr()), "euler_step")
            tr.launches += 1
            sched.advance()
            if callback_on_step_end is not None:
                out = callback_on_step_end(self, i, t, {"latents": latents, "prompt_embeds": prompt_embeds})
                latents = out.pop("latents", latents)
                prompt_embeds = out.pop("prompt_embeds", prompt_embeds)
        return latents

    @torch.no_grad()
    def __call__(
        self,
        prompt=None,
        prompt_2=None,
        prompt_embeds: torch.Tensor | None = None,
        pooled_prompt_embeds: torch.Tensor | None = None,
        num_images_per_prompt: int = 1,
        num_inference_steps: int = 20,
        generator: torch.Generator | None = None,
        latents: torch.Tensor | None = None,
        guidance_scale: float = 5.0,
        height: int = 256,
        width: int = 256,
        callback_on_step_end: Callable[..., dict[str, torch.Tensor]] | None = None,
        callback_on_step_end_tensor_inputs: list[str] | None = None,
        output_type: str = "latent",
        return_dict: bool = False,
        capture_callback: Callable[..., dict[str, torch.Tensor]] | None = None,
        **kwargs: Any,
    ):
        if prompt is not None or prompt_2 is not None:
            raise ValueError("text encoding is out of scope: pass prompt_embeds / pooled_prompt_embeds")
        if prompt_embeds is None or pooled_prompt_embeds is None:
            raise ValueError("prompt_embeds and pooled_prompt_embeds are required")
        if output_type != "latent":
            raise NotImplementedError("VAE decode is out of scope; use output_type='latent'")
        if num_images_per_prompt != 1:
            raise NotImplementedError("the reference always calls with num_images_per_prompt=1")
        tr, dev, sched = self.transformer, self.device, self.scheduler
        batch_size = prompt_embeds.shape[0]
        prompt_embeds = prompt_embeds.to(dev)
        pooled_prompt_embeds = pooled_prompt_embeds.to(dev)
        text_ids = torch.zeros(batch_size, prompt_embeds.shape[1], 3)
        latents, img_ids = self.prepare_latents(batch_size, tr.config.in_channels // 4, height, width, generator, latents)
        n_tokens = latents.shape[1]
        sc = sched.config
        mu = calculate_shift(n_tokens, sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift)
        sched.set_timesteps(num_inference_steps, device=dev, mu=mu)
        guidance = torch.full((batch_size,), float(guidance_scale), dtype=torch.float32, device=dev) \
            if tr.config.guidance_embeds else None
        inputs = {"latents": latents.contiguous(), "prompt_embeds": prompt_embeds,
                  "pooled_prompt_embeds": pooled_prompt_embeds, "guidance": guidance,
                  "timesteps": (sched.timesteps / 1000).to(dev)}
        if self.use_cuda_graph:
            # keyed on the schedule's CONTENT (not its id/name: a freed candidate's address can be reused) and on the
            # transformer's buffer epoch (a recorded graph holds raw pointers into its workspace)
            if getattr(tr, "buffer_epoch", 0) != getattr(self, "_graph_epoch", None):
                self._graphs.clear()
                self._graph_epoch = getattr(tr, "buffer_epoch", 0)
            key = ("flux", tr.cache_schedule.content_key(), tuple(latents.shape), tuple(prompt_embeds.shape),
                   num_inference_steps, float(guidance_scale), float(mu))
            latents = self._graphs.run(
                key, inputs, lambda st, cb: self._denoise(st, text_ids, img_ids, cb),
                capture_callback if capture_callback is not None else callback_on_step_end, tr)
            if getattr(tr, "buffer_epoch", 0) != self._graph_epoch:
                # the run's own eager warm-up re-allocated the workspace / per-timestep tables: the graph just recorded
                # points into the new buffers, every OLDER graph into freed memory
                self._graphs.keep_only(key)
                self._graph_epoch = getattr(tr, "buffer_epoch", 0)
            if callback_on_step_end is not None:
                for i, t in enumerate(sched.timesteps):
                    out = callback_on_step_end(self, i, t, {"latents": latents, "prompt_embeds": prompt_embeds})
                    latents = out.pop("latents", latents)
        else:
            latents = self._denoise(inputs, text_ids, img_ids, callback_on_step_end)
        if not return_dict:
            return (latents,)
        return {"images": latents}
