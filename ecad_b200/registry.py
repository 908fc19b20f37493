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
    """custom_attn_ff.py:6-49 - register by lower-cased function name; unknown/None name -> DEFAULT."""

    _registry: dict[str, Callable[[DecisionContext], bool]] = {}
    DEFAULT: str = ""

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
    DEFAULT = "compute_attn_cached"


class ComputeFFRegistry(ComputeRegistry):
    _registry: dict[str, Callable[[DecisionContext], bool]] = {}
    DEFAULT = "compute_ff_cached"


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
    def get(cls, name: str) -> type:
        if name not in cls.registry:
            raise ValueError(f"Image generator {name} not found. Available: {sorted(cls.registry)}")
        return cls.registry[name]
