"""Re-keying of diffusers-named state dicts into the names of the independent implementations the anchor tests
compare against (tests/test_third_party_anchors.py on the CPU, tests/test_gpu_anchors.py on the GPU): the inverse of
diffusers' published conversion scripts (scripts/convert_flux_to_diffusers.py, convert_ldm_vae_checkpoint)."""
import torch


def bfl_state_dict(sd, L, LS, D):
    """diffusers FluxTransformer2DModel keys -> BFL keys (inverse of diffusers' convert_flux_to_diffusers.py)."""
    out = {}

    def lin(dst, src):
        out[dst + ".weight"], out[dst + ".bias"] = sd[src + ".weight"], sd[src + ".bias"]

    def cat(dst, srcs):
        out[dst + ".weight"] = torch.cat([sd[s + ".weight"] for s in srcs], 0)
        out[dst + ".bias"] = torch.cat([sd[s + ".bias"] for s in srcs], 0)

    lin("img_in", "x_embedder")
    lin("txt_in", "context_embedder")
    lin("time_in.in_layer", "time_text_embed.timestep_embedder.linear_1")
    lin("time_in.out_layer", "time_text_embed.timestep_embedder.linear_2")
    lin("vector_in.in_layer", "time_text_embed.text_embedder.linear_1")
    lin("vector_in.out_layer", "time_text_embed.text_embedder.linear_2")
    for i in range(L):
        s, d = f"transformer_blocks.{i}", f"double_blocks.{i}"
        lin(f"{d}.img_mod.lin", f"{s}.norm1.linear")
        lin(f"{d}.txt_mod.lin", f"{s}.norm1_context.linear")
        cat(f"{d}.img_attn.qkv", [f"{s}.attn.to_q", f"{s}.attn.to_k", f"{s}.attn.to_v"])
        cat(f"{d}.txt_attn.qkv", [f"{s}.attn.add_q_proj", f"{s}.attn.add_k_proj", f"{s}.attn.add_v_proj"])
        out[f"{d}.img_attn.norm.query_norm.weight"] = sd[f"{s}.attn.norm_q.weight"]
        out[f"{d}.img_attn.norm.key_norm.weight"] = sd[f"{s}.attn.norm_k.weight"]
        out[f"{d}.txt_attn.norm.query_norm.weight"] = sd[f"{s}.attn.norm_added_q.weight"]
        out[f"{d}.txt_attn.norm.key_norm.weight"] = sd[f"{s}.attn.norm_added_k.weight"]
        lin(f"{d}.img_attn.proj", f"{s}.attn.to_out.0")
        lin(f"{d}.txt_attn.proj", f"{s}.attn.to_add_out")
        lin(f"{d}.img_mlp.0", f"{s}.ff.net.0.proj")
        lin(f"{d}.img_mlp.2", f"{s}.ff.net.2")
        lin(f"{d}.txt_mlp.0", f"{s}.ff_context.net.0.proj")
        lin(f"{d}.txt_mlp.2", f"{s}.ff_context.net.2")
    for i in range(LS):
        s, d = f"single_transformer_blocks.{i}", f"single_blocks.{i}"
        lin(f"{d}.modulation.lin", f"{s}.norm.linear")
        cat(f"{d}.linear1", [f"{s}.attn.to_q", f"{s}.attn.to_k", f"{s}.attn.to_v", f"{s}.proj_mlp"])
        lin(f"{d}.linear2", f"{s}.proj_out")
        out[f"{d}.norm.query_norm.weight"] = sd[f"{s}.attn.norm_q.weight"]
        out[f"{d}.norm.key_norm.weight"] = sd[f"{s}.attn.norm_k.weight"]
    lin("final_layer.linear", "proj_out")
    # AdaLayerNormContinuous chunks (scale, shift); BFL's LastLayer chunks (shift, scale): the halves swap
    w, b = sd["norm_out.linear.weight"], sd["norm_out.linear.bias"]
    out["final_layer.adaLN_modulation.1.weight"] = torch.cat([w[D:], w[:D]], 0)
    out["final_layer.adaLN_modulation.1.bias"] = torch.cat([b[D:], b[:D]], 0)
    return out


def ldm_decoder_state_dict(sd, n_up: int = 4, layers: int = 3):
    """diffusers AutoencoderKL decoder keys -> ldm Decoder keys (mid.block_1 / attn_1 / block_2, up.{i} counted from
    the IMAGE side, nin_shortcut, 1x1-convolution attention projections)."""
    out = {}

    def copy(dst, src):
        out[dst + ".weight"], out[dst + ".bias"] = sd[src + ".weight"], sd[src + ".bias"]

    def resnet(dst, src):
        for n in ("norm1", "conv1", "norm2", "conv2"):
            copy(f"{dst}.{n}", f"{src}.{n}")
        if src + ".conv_shortcut.weight" in sd:
            copy(f"{dst}.nin_shortcut", f"{src}.conv_shortcut")

    copy("conv_in", "decoder.conv_in")
    resnet("mid.block_1", "decoder.mid_block.resnets.0")
    resnet("mid.block_2", "decoder.mid_block.resnets.1")
    copy("mid.attn_1.norm", "decoder.mid_block.attentions.0.group_norm")
    for dst, src in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):  # 1x1 convs <-> linears
        out[f"mid.attn_1.{dst}.weight"] = sd[f"decoder.mid_block.attentions.0.{src}.weight"][:, :, None, None]
        out[f"mid.attn_1.{dst}.bias"] = sd[f"decoder.mid_block.attentions.0.{src}.bias"]
    for i in range(n_up):  # diffusers counts the up blocks from the latent side, ldm from the image side
        for j in range(layers):
            resnet(f"up.{n_up - 1 - i}.block.{j}", f"decoder.up_blocks.{i}.resnets.{j}")
        if i < n_up - 1:
            copy(f"up.{n_up - 1 - i}.upsample.conv", f"decoder.up_blocks.{i}.upsamplers.0.conv")
    copy("norm_out", "decoder.conv_norm_out")
    copy("conv_out", "decoder.conv_out")
    return out
