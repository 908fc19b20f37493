"""String-keyed plug-in registries, mirroring the reference's (same names, same lookup rules).

Reference: /root/reference/ecad/transformer_blocks/custom_attn_ff.py:6-59 (ComputeRegistry / ComputeAttnRegistry /
ComputeFFRegistry) and /root/reference/ecad/image_generators/load_image_generator.py:16-84 (ImageGeneratorRegistry).

On the B200 path a "compute function" is not a Python callable over tensors (the sub-blocks run inside the C
executor); what is registered is the *decision policy* the host applies before it hands the executed-mask to
`ecadk_pixart_blocks`.  Policies have the signature ``policy(ctx: DecisionContext) -> bool`` ("run the module?").
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable


@dataclass
class DecisionContext:
    """What a policy may look at: mirrors the state `compute_attn_cached` reads in the reference."""

    block: int
    component: str  # "attn1" | "attn2" | "ff"
    recompute: bool  # schedule[curr_step][block][component]
    no_cache: bool  # the cached tensor is None
    curr_step: int
    kwargs: dict[str, Any]


class ComputeRegistry:
    """custom_attn_ff.py:6-49 - register by lower-cased function name; unknown/None name -> DEFAULT.

    Two kinds of entries share the name space:
      * decision policies ``policy(ctx) -> bool`` (``register``): the fast path - the host only decides run / reuse and
        the C executor does the rest;
      * tensor functions with the REFERENCE's signature (``register_tensor``):
        ``f(block, attn, hidden_states, encoder_hidden_states, attention_mask, **kwargs) -> Tensor`` for attention and
        ``f(block, norm_hidden_states, **kwargs) -> Tensor`` for the feed-forward
        (cached_transformer_block.py:141-149,161-165).  ``block`` is a `B200BlockProxy`: ``block.attn1`` / ``attn2`` /
        ``ff`` run the sm_100a kernels of that block, ``block.cached_attn1_output`` etc. are the HBM cache slots,
        ``block.cache_schedule`` / ``block.block_num`` as in the reference.  A block that names a tensor function at a
        step is executed sub-block by sub-block from Python at that step (the other blocks stay in the C executor).
    """

    _registry: dict[str, Callable[[DecisionContext], bool]] = {}
    _tensor_registry: dict[str, Callable[..., Any]] = {}
    DEFAULT: str = ""

    @classmethod
    def register_tensor(cls, func: Callable[..., Any]) -> Callable[..., Any]:
        """Decorator with the reference's registration rule (custom_attn_ff.py:10-20: lower-cased ``__name__``)."""
        name = func.__name__.lower()
        if name in cls._registry:
            raise ValueError(f"'{name}' is a built-in decision policy; pick another name for a tensor function")
        cls._tensor_registry[name] = func
        return func

    @classmethod
    def get_tensor(cls, key: str | None):
        return cls._tensor_registry.get(key.lower()) if key is not None else None

    @classmethod
    def register(cls, func: Callable[[DecisionContext], bool]) -> Callable[[DecisionContext], bool]:
        cls._registry[func.__name__.lower()] = func
        return func

    @classmethod
    def get(cls, key: str | None, none_if_not_found: bool = False):
        if key is not None:
            key = key.lower()
            if key in cls._registry:
                return cls._registry[key]
        return None if none_if_not_found else cls.default()

    @classmethod
    def default(cls):
        if not cls.DEFAULT:
            raise NotImplementedError("Subclasses must define a DEFAULT attribute.")
        func = cls._registry.get(cls.DEFAULT)
        if func is None:
            raise ValueError(f"Default function '{cls.DEFAULT}' not registered.")
        return func


class ComputeAttnRegistry(ComputeRegistry):
    _registry: dict[str, Callable[[DecisionContext], bool]] = {}
    _tensor_registry: dict[str, Callable[..., Any]] = {}
    DEFAULT = "compute_attn_cached"


class ComputeFFRegistry(ComputeRegistry):
    _registry: dict[str, Callable[[DecisionContext], bool]] = {}
    _tensor_registry: dict[str, Callable[..., Any]] = {}
    DEFAULT = "compute_ff_cached"


# Tensor-level restatements of the reference defaults, for use INSIDE user functions and for the sub-blocks of a
# Python-executed block that keep the default behaviour (cached_transformer_block.py:326-391).
def compute_attn_cached_tensor(block, attn: str, hidden_states, encoder_hidden_states=None, attention_mask=None,
                               **cross_attention_kwargs):
    if attn not in ("attn1", "attn2"):
        raise ValueError(f"Invalid attention type: {attn}. Must be attn1 or attn2")
    recompute = block.cache_schedule.get_recompute(block.block_num, attn)
    cached = getattr(block, f"cached_{attn}_output")
    if recompute or cached is None:
        if not recompute:
            print(f"WARNING: No cached {attn} found. Recomputing.")
        out = getattr(block, attn)(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                   attention_mask=attention_mask, **cross_attention_kwargs)
    else:
        out = cached
    setattr(block, f"cached_{attn}_output", out)
    return out


def compute_ff_cached_tensor(block, norm_hidden_states, **kwargs):
    recompute = block.cache_schedule.get_recompute(block.block_num, "ff")
    cached = block.cached_ff_output
    if recompute or cached is None:
        if not recompute:
            print("WARNING: No cached ff found. Recomputing.")
        out = block.ff(norm_hidden_states)
    else:
        out = cached
    block.cached_ff_output = out
    return out


def _warn_no_cache(ctx: DecisionContext) -> None:
    # cached_transformer_block.py:344-345, :370-371 - warn and fall back to recompute
    if not ctx.recompute and ctx.no_cache:
        print(f"WARNING: No cached {ctx.component} found. Recomputing.")


@ComputeAttnRegistry.register
def compute_attn_cached(ctx: DecisionContext) -> bool:
    """cached_transformer_block.py:326-360: run iff ``recompute or cache is None``."""
    if ctx.component not in ("attn1", "attn2"):
        raise ValueError(f"Invalid attention type: {ctx.component}. Must be attn1 or attn2")
    _warn_no_cache(ctx)
    return ctx.recompute or ctx.no_cache


@ComputeFFRegistry.register
def compute_ff_cached(ctx: DecisionContext) -> bool:
    """cached_transformer_block.py:362-391."""
    _warn_no_cache(ctx)
    return ctx.recompute or ctx.no_cache


@ComputeAttnRegistry.register
def compute_attn_tgate(ctx: DecisionContext) -> bool:
    """cached_transformer_block.py:393-454: attn1 as usual; attn2 follows the flag rule while
    ``curr_step <= gate_step - 1`` and is never run from ``gate_step`` on."""
    if ctx.component not in ("attn1", "attn2"):
        raise ValueError(f"Invalid attention type: {ctx.component}. Must be attn1 or attn2")
    gate_step = ctx.kwargs.get("gate_step")
    if gate_step is None:
        raise ValueError("gate_step must be provided as a kwarg to commpute_attn_tgate.")
    if ctx.component == "attn1" or ctx.curr_step <= gate_step - 1:
        return compute_attn_cached(ctx)
    assert not ctx.no_cache, "Cross-Attention must be cached at gate step for TGATE."
    return False


class ImageGeneratorRegistry:
    """load_image_generator.py:16-84 - name -> ImageGenerator class."""

    registry: dict[str, type] = {}

    @classmethod
    def register(cls, name: str):
        def deco(klass: type) -> type:
            cls.registry[name] = klass
            return klass

        return deco

    @classmethod
    def get(cls, name: str, default_name: str | None = None) -> type | None:
        """load_image_generator.py:23-40: the class registered under ``name``, else the one under ``default_name``,
        else ``None`` (the callers raise, see get_image_generator_type)."""
        klass = cls.registry.get(name, None)
        if klass is None and default_name is not None:
            klass = cls.registry.get(default_name, None)
        return klass


def get_image_generator_type_from_config(config: dict, default_name: str = "PixArtImageGenerator") -> type:
    """load_image_generator.py:43-66."""
    klass = None
    if "image_generator" in config:
        klass = ImageGeneratorRegistry.get(config["image_generator"], default_name)
    if klass is None:
        raise ValueError(f"Image generator not found in config: {config}.")
    return klass


def get_image_generator_type(name: str, default_name: str = "PixArtImageGenerator") -> type:
    """load_image_generator.py:69-84."""
    klass = ImageGeneratorRegistry.get(name, default_name)
    if klass is None:
        raise ValueError(f"Image generator not found: {name}.")
    return klass


class PipelineRegistry:
    """/root/reference/ecad/pipelines/load_pipeline.py:16-41 - the third string-keyed plug-in point (SURVEY.md section
    5): ``config.pipeline.name`` of a schedule JSON -> pipeline class, under the reference's own names.  The two stock
    diffusers loops (``pixart_alpha`` / ``pixart_sigma``) and ``pass_through`` (a verbatim copy of the same loop,
    pass_through.py:186-404) are one class here; ``tgate`` adds the gate step; ``flux`` is the flow-match loop.
    ``register`` lets a user add a pipeline class of their own (anything with
    ``from_pretrained(transformer, **kwargs)``)."""

    _registry: dict[str, type] = {}
    _builtin_loaded = False

    @classmethod
    def _builtin(cls) -> dict[str, type]:
        if not cls._builtin_loaded:  # resolved lazily: the pipeline modules import this one
            from .flux_pipeline import B200FluxPipeline
            from .pipeline import B200PixArtPipeline, B200TGATEPipeline

            for name, klass in (("pixart_alpha", B200PixArtPipeline), ("pixart_sigma", B200PixArtPipeline),
                                ("tgate", B200TGATEPipeline), ("flux", B200FluxPipeline),
                                ("pass_through", B200PixArtPipeline)):
                cls._registry.setdefault(name, klass)
            cls._builtin_loaded = True
        return cls._registry

    @classmethod
    def register(cls, name: str):
        def deco(klass: type) -> type:
            cls._builtin()[name] = klass
            return klass

        return deco

    @classmethod
    def get(cls, name: str | None, default_name: str | None = None) -> type | None:
        """load_pipeline.py:25-41: the class under ``name``, else the one under ``default_name``, else ``None``."""
        reg = cls._builtin()
        klass = reg.get(name, None)
        if klass is None and default_name is not None:
            klass = reg.get(default_name, None)
        return klass


def pipeline_from_pretrained(pipeline_config: dict | None, default_pipeline_name: str | None = None):
    """load_pipeline.py:44-58: resolve ``{"name": ..., "kwargs": {...}}`` to a factory that forwards the extra kwargs
    (TGATE's ``gate_step``).  As in the reference the default name applies when the config has no ``name`` key; an
    unknown name is an error here (the reference fails one line later, calling ``None.from_pretrained``)."""
    pipeline_config = pipeline_config or {}
    name = pipeline_config.get("name", default_pipeline_name)
    klass = PipelineRegistry.get(name)
    if klass is None:
        raise ValueError(f"Pipeline not found: {name!r}.")
    extra_kwargs = dict(pipeline_config.get("kwargs") or {})

    def from_pretrained(*args, **kwargs):
        return klass.from_pretrained(*args, **kwargs, **extra_kwargs)

    from_pretrained.pipeline_class = klass
    from_pretrained.extra_kwargs = extra_kwargs
    return from_pretrained
