"""Runtime that replaces ``unet.forward``: per-(batch, H, W, dtype) plan cache + CUDA-graph replay.

Reference counterparts: the lazy-trace wrapper (/root/reference/src/sfast/jit/trace_helper.py:
33-72) and the dynamic CUDA-graph cache (/root/reference/src/sfast/cuda/graphs.py:16-51,
147-157).  Same contract: inputs are copied into static buffers, the graph is replayed, the
output is returned as a fresh clone, all under a per-object lock; unlike the reference the
timestep is always read from device memory inside the graph (never keyed by value,
graphs.py:229-231).
"""
import logging
import threading
import time

import torch

from .plan import PackedWeights, UNetPlan
from .svd_plan import SVDPlan
from .unet_spec import cfg_get, spec_from_config

logger = logging.getLogger(__name__)

try:  # diffusers is not a dependency; use its output type when present
    from diffusers.models.unets.unet_2d_condition import UNet2DConditionOutput  # type: ignore
except Exception:  # noqa: BLE001
    class UNet2DConditionOutput(dict):
        """Minimal stand-in for diffusers' BaseOutput: attribute, key and index access."""

        def __init__(self, sample):
            super().__init__(sample=sample)
            self.sample = sample

        def __getitem__(self, k):
            if isinstance(k, int):
                return tuple(self.values())[k]
            return super().__getitem__(k)

        def to_tuple(self):
            return tuple(self.values())


_UNSUPPORTED_KWARGS = ("class_labels", "timestep_cond", "attention_mask",
                       "down_intrablock_additional_residuals", "encoder_attention_mask")


def require_b200(device):
    if device.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("sfast (B200 build): the UNet hot path needs a CUDA sm_100 device; "
                           "there is no CPU / eager fallback")
    major, minor = torch.cuda.get_device_capability(device)
    if (major, minor) != (10, 0):  # the cubin is sm_100a: no forward compatibility to sm_103 / sm_110
        raise RuntimeError(f"sfast (B200 build): kernels are sm_100a only, device is sm_{major}{minor}")


class _GraphedPlan:
    def __init__(self, plan: UNetPlan, use_graph: bool):
        self.plan = plan
        self.graph = None
        stream = torch.cuda.current_stream()
        # warm-up (also sets kernel attributes, which must not happen during capture)
        side = torch.cuda.Stream()
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            plan.run()
        stream.wait_stream(side)
        torch.cuda.synchronize()
        if use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                plan.run()
            self.graph = g

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.plan.run()


class CompiledUNet:
    """Callable with the signature of diffusers ``UNet2DConditionModel.forward``."""

    def __init__(self, config, state_dict_fn, enable_cuda_graph=True, preserve_parameters=True):
        self._config = config
        self._state_dict_fn = state_dict_fn
        self.enable_cuda_graph = enable_cuda_graph
        # reference contract (preserve_parameters=True): an in-place parameter update is seen by
        # the next call.  The packed weights are copies, so the parameters' version counters are
        # compared on every call and the copies refreshed in place when they moved.
        self.preserve_parameters = preserve_parameters
        self._weights = None
        self._param_refs = None
        self._param_versions = None
        self._cached = {}
        self._lock = threading.Lock()
        self.spec = spec_from_config(config)
        self.first_call_s = None  # wall time of the first call (packing + plan + capture)

    # -- weights -------------------------------------------------------------------------
    def _versions(self):
        return [p._version for p in self._param_refs]

    def _ensure_weights(self, dtype, device):
        if self._weights is None or self._weights.dtype != dtype or self._weights.device != device:
            logger.info("Packing UNet weights for the B200 path (%s, %s)", dtype, device)
            sd = self._state_dict_fn()
            self._weights = PackedWeights(self.spec, sd, dtype, device)
            self._param_refs = [t for t in sd.values() if torch.is_tensor(t)]
            self._param_versions = self._versions()
            self._cached.clear()
        elif self.preserve_parameters:
            now = self._versions()
            if now != self._param_versions:
                logger.info("UNet parameters changed in place: refreshing the packed weights")
                self._refresh_locked()
        return self._weights

    def _refresh_locked(self):
        sd = self._state_dict_fn()
        torch.cuda.current_stream().synchronize()  # no replay may be reading the old values
        self._weights.refresh(sd)
        self._param_refs = [t for t in sd.values() if torch.is_tensor(t)]
        self._param_versions = self._versions()

    def rebind(self):
        """Re-pack the weights now (same device storage: plans and CUDA graphs stay valid).
        Only needed with preserve_parameters=False or after replacing parameter objects in a way
        the version counters cannot see."""
        with self._lock:
            if self._weights is not None:
                with torch.cuda.device(self._weights.device):
                    self._refresh_locked()

    # -- forward -------------------------------------------------------------------------
    def _call_svd(self, sample, timestep, encoder_hidden_states, added_time_ids, return_dict=True):
        """diffusers UNetSpatioTemporalConditionModel.forward(sample [B, F, C, H, W], timestep,
        encoder_hidden_states [B, 1, ctx], added_time_ids [B, 3])."""
        require_b200(sample.device)
        dtype = sample.dtype
        if dtype not in (torch.float16, torch.bfloat16):
            raise NotImplementedError(f"sfast (B200 build): UNet dtype {dtype}; use fp16 or bf16")
        B, F_, _, H, W = sample.shape
        if F_ != self.spec.num_frames:
            raise NotImplementedError(f"sfast (B200 build): {F_} frames, the UNet was configured for "
                                      f"{self.spec.num_frames}")
        if encoder_hidden_states.shape[1] != 1:
            raise NotImplementedError("sfast (B200 build): SVD path needs a single-token encoder_hidden_states")
        with self._lock, torch.cuda.device(sample.device):
            t_first = time.perf_counter() if self.first_call_s is None else None
            weights = self._ensure_weights(dtype, sample.device)
            key = ("svd", B, H, W, dtype, sample.device.index)
            gp = self._cached.get(key)
            if gp is None:
                logger.info("Building SVD UNet launch plan for %s", key)
                gp = _GraphedPlan(SVDPlan(weights, B, H, W), self.enable_cuda_graph)
                self._cached[key] = gp
            plan = gp.plan
            plan.sample_in.copy_(sample.reshape(B * F_, -1, H, W), non_blocking=True)
            plan.ehs_in.copy_(encoder_hidden_states.repeat_interleave(F_, dim=0), non_blocking=True)
            t = timestep if torch.is_tensor(timestep) else torch.tensor(float(timestep))
            t = t.reshape(-1).to(device=sample.device, dtype=torch.float32, non_blocking=True)
            plan.t_in.copy_((t.expand(B) if t.numel() == 1 else t).repeat_interleave(F_))
            plan.time_ids_in.copy_(added_time_ids.to(device=sample.device, dtype=torch.float32)
                                   .repeat_interleave(F_, dim=0).reshape(-1))
            gp.step()
            out = plan.out.clone().reshape(B, F_, -1, H, W)
            if t_first is not None:
                torch.cuda.current_stream().synchronize()
                self.first_call_s = time.perf_counter() - t_first
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(sample=out)

    def __call__(self, sample, timestep, encoder_hidden_states, *args, **kwargs):
        if self.spec.temporal:
            return self._call_svd(sample, timestep, encoder_hidden_states, *args, **kwargs)
        return self._call_2d(sample, timestep, encoder_hidden_states, *args, **kwargs)

    def _call_2d(self, sample, timestep, encoder_hidden_states, class_labels=None,
                 timestep_cond=None, attention_mask=None, cross_attention_kwargs=None,
                 added_cond_kwargs=None, down_block_additional_residuals=None,
                 mid_block_additional_residual=None, down_intrablock_additional_residuals=None,
                 encoder_attention_mask=None, return_dict=True):
        given = dict(class_labels=class_labels, timestep_cond=timestep_cond,
                     attention_mask=attention_mask,
                     down_intrablock_additional_residuals=down_intrablock_additional_residuals,
                     encoder_attention_mask=encoder_attention_mask)
        for k in _UNSUPPORTED_KWARGS:
            if given[k] is not None:
                raise NotImplementedError(f"sfast (B200 build): UNet argument `{k}` is not supported")
        if cross_attention_kwargs:
            scale = cross_attention_kwargs.get("scale", 1.0)
            if set(cross_attention_kwargs) - {"scale"} or scale != 1.0:
                raise NotImplementedError("sfast (B200 build): cross_attention_kwargs (LoRA scale) "
                                          "is not supported; fuse the LoRA weights into the parameters")
        require_b200(sample.device)
        dtype = sample.dtype
        if dtype not in (torch.float16, torch.bfloat16):
            raise NotImplementedError(f"sfast (B200 build): UNet dtype {dtype}; use fp16 or bf16")
        B, _, H, W = sample.shape
        ctx_len = encoder_hidden_states.shape[1]
        controlnet = down_block_additional_residuals is not None or mid_block_additional_residual is not None
        if controlnet and (down_block_additional_residuals is None or mid_block_additional_residual is None):
            raise NotImplementedError("sfast (B200 build): ControlNet residuals need both "
                                      "down_block_additional_residuals and mid_block_additional_residual")
        # everything below (allocation, TMA maps, launches, capture) targets sample's device
        with self._lock, torch.cuda.device(sample.device):
            t_first = time.perf_counter() if self.first_call_s is None else None
            weights = self._ensure_weights(dtype, sample.device)
            key = (B, H, W, dtype, ctx_len, sample.device.index, controlnet)
            gp = self._cached.get(key)
            if gp is None:
                logger.info("Building UNet launch plan for %s", key)
                plan = UNetPlan(weights, B, H, W, ctx_len, controlnet=controlnet)
                gp = _GraphedPlan(plan, self.enable_cuda_graph)
                self._cached[key] = gp
            plan = gp.plan
            plan.sample_in.copy_(sample, non_blocking=True)
            plan.ehs_in.copy_(encoder_hidden_states, non_blocking=True)
            t = timestep if torch.is_tensor(timestep) else torch.tensor(float(timestep))
            t = t.reshape(-1).to(device=sample.device, dtype=torch.float32, non_blocking=True)
            plan.t_in.copy_(t.expand(B) if t.numel() == 1 else t)
            if self.spec.addition_embed_type == "text_time":
                if not added_cond_kwargs or "text_embeds" not in added_cond_kwargs:
                    raise ValueError("added_cond_kwargs with text_embeds / time_ids is required")
                te = added_cond_kwargs["text_embeds"]
                plan.add_in[:, :te.shape[1]].copy_(te)
                plan.time_ids_in.copy_(added_cond_kwargs["time_ids"].reshape(-1).float())
            if controlnet:
                res = list(down_block_additional_residuals) + [mid_block_additional_residual]
                if len(res) != len(plan.ctrl_in):
                    raise ValueError(f"expected {len(plan.ctrl_in) - 1} down-block residuals, got {len(res) - 1}")
                for dst, src in zip(plan.ctrl_in, res):
                    if tuple(src.shape) != tuple(dst.shape):
                        raise ValueError(f"ControlNet residual shape {tuple(src.shape)} != {tuple(dst.shape)}")
                    dst.copy_(src, non_blocking=True)
            gp.step()
            out = plan.out.clone()
            if t_first is not None:
                torch.cuda.current_stream().synchronize()
                self.first_call_s = time.perf_counter() - t_first
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(sample=out)


class DecoderOutput(dict):
    """Stand-in for diffusers' DecoderOutput (attribute, key and index access)."""

    def __init__(self, sample):
        super().__init__(sample=sample)
        self.sample = sample

    def __getitem__(self, k):
        return tuple(self.values())[k] if isinstance(k, int) else super().__getitem__(k)


class CompiledVAEDecoder:
    """`AutoencoderKL.decode(z, return_dict=True, generator=None)` on the native path."""

    def __init__(self, config, state_dict_fn, eager_decode, enable_cuda_graph=True):
        from .vae_plan import vae_spec_from_config
        self.spec = vae_spec_from_config(config)
        self._state_dict_fn, self._eager = state_dict_fn, eager_decode
        self.enable_cuda_graph = enable_cuda_graph
        self._weights, self._param_refs, self._param_versions = None, None, None
        self._cached = {}
        self._lock = threading.Lock()
        self._warned = False

    def __call__(self, z, return_dict=True, generator=None):
        from .vae_plan import VAEDecodePlan
        if z.dtype not in (torch.float16, torch.bfloat16):
            # e.g. SDXL's force_upcast fp32 VAE: the whole module stays on its own eager path
            if not self._warned:
                logger.warning("sfast (B200 build): VAE decode in %s is left on the module's eager path "
                               "(the native decoder computes in fp16 / bf16)", z.dtype)
                self._warned = True
            return self._eager(z, return_dict=return_dict, generator=generator)
        require_b200(z.device)
        B, _, H, W = z.shape
        with self._lock, torch.cuda.device(z.device):
            if self._weights is None or self._weights.dtype != z.dtype or self._weights.device != z.device:
                sd = self._state_dict_fn()
                self._weights = PackedWeights(self.spec, sd, z.dtype, z.device)
                self._param_refs = [t for t in sd.values() if torch.is_tensor(t)]
                self._param_versions = [t._version for t in self._param_refs]
                self._cached.clear()
            elif [t._version for t in self._param_refs] != self._param_versions:
                sd = self._state_dict_fn()
                torch.cuda.current_stream().synchronize()
                self._weights.refresh(sd)
                self._param_refs = [t for t in sd.values() if torch.is_tensor(t)]
                self._param_versions = [t._version for t in self._param_refs]
            key = (B, H, W, z.dtype, z.device.index)
            gp = self._cached.get(key)
            if gp is None:
                gp = _GraphedPlan(VAEDecodePlan(self._weights, B, H, W), self.enable_cuda_graph)
                self._cached[key] = gp
            gp.plan.z_in.copy_(z, non_blocking=True)
            gp.step()
            out = gp.plan.out.clone()
        if not return_dict:
            return (out,)
        try:
            from diffusers.models.autoencoders.vae import DecoderOutput as DO  # type: ignore
            return DO(sample=out)
        except Exception:  # noqa: BLE001
            return DecoderOutput(sample=out)


def compile_vae_module(m, enable_cuda_graph=True):
    """Replace ``m.decode`` of an AutoencoderKL (same module object; `encode` is left alone)."""
    eager = m.decode
    compiled = CompiledVAEDecoder(m.config, m.state_dict, eager, enable_cuda_graph)

    def decode(z, return_dict=True, generator=None):
        return compiled(z, return_dict=return_dict, generator=generator)

    decode.__self__ = m
    decode._cached = compiled._cached
    decode._compiled = compiled
    m.decode = decode
    return m


class CompiledTextEncoder:
    """transformers ``CLIPTextModel`` / ``CLIPTextModelWithProjection`` forward on the native path: the
    text encoders the reference traces and graphs
    (/root/reference/src/sfast/compilers/diffusion_pipeline_compiler.py:93-112).  One plan + CUDA graph
    per (batch, sequence length); outputs are fresh tensors."""

    def __init__(self, config, state_dict_fn, eager_forward, with_projection, enable_cuda_graph=True):
        from .clip_plan import clip_text_spec_from_config
        self.spec = clip_text_spec_from_config(config, with_projection)
        self._state_dict_fn, self._eager = state_dict_fn, eager_forward
        self.with_projection = with_projection
        self.enable_cuda_graph = enable_cuda_graph
        self._weights, self._param_refs, self._param_versions = None, None, None
        self._cached = {}
        self._lock = threading.Lock()
        self._warned = False
        rd = cfg_get(config, "return_dict", None)
        self._return_dict_default = True if rd is None else bool(rd)

    def _route_eager(self, why, args, kwargs):
        if not self._warned:
            logger.warning("sfast (B200 build): text encoder call left on the module's eager path (%s)", why)
            self._warned = True
        return self._eager(*args, **kwargs)

    def __call__(self, input_ids=None, attention_mask=None, position_ids=None, output_attentions=None,
                 output_hidden_states=None, return_dict=None, **kwargs):
        from .clip_plan import ClipTextPlan
        call = dict(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                    output_attentions=output_attentions, output_hidden_states=output_hidden_states,
                    return_dict=return_dict, **kwargs)
        call = {k: v for k, v in call.items() if v is not None}
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        if attention_mask is not None or position_ids is not None or output_attentions or kwargs:
            return self._route_eager("attention_mask / position_ids / output_attentions / extra kwargs", (), call)
        sd = None
        if self._weights is None:
            sd = self._state_dict_fn()
            wdt = sd["text_model.embeddings.token_embedding.weight"].dtype
            if wdt not in (torch.float16, torch.bfloat16):
                return self._route_eager(f"{wdt} weights: the native encoder computes in fp16 / bf16", (), call)
        ids = input_ids.view(-1, input_ids.shape[-1])
        require_b200(ids.device)
        B, S = ids.shape
        with self._lock, torch.cuda.device(ids.device):
            if self._weights is None:
                wdt = sd["text_model.embeddings.token_embedding.weight"].dtype
                self._weights = PackedWeights(self.spec, sd, wdt, ids.device)
                self._param_refs = [t for t in sd.values() if torch.is_tensor(t)]
                self._param_versions = [t._version for t in self._param_refs]
            elif [t._version for t in self._param_refs] != self._param_versions:
                sd = self._state_dict_fn()
                torch.cuda.current_stream().synchronize()
                self._weights.refresh(sd)
                self._param_refs = [t for t in sd.values() if torch.is_tensor(t)]
                self._param_versions = [t._version for t in self._param_refs]
            key = (B, S, ids.device.index)
            gp = self._cached.get(key)
            if gp is None:
                gp = _GraphedPlan(ClipTextPlan(self._weights, B, S), self.enable_cuda_graph)
                self._cached[key] = gp
            plan = gp.plan
            plan.ids_in.copy_(ids, non_blocking=True)
            gp.step()
            last = plan.last_hidden_state.clone()
            pooled = plan.pooled.clone()
            embeds = plan.text_embeds.clone() if self.with_projection else None
            hidden = tuple(h.clone() for h in plan.hidden_states()) if output_hidden_states else None
        if return_dict is None:
            return_dict = self._return_dict_default
        if self.with_projection:
            if not return_dict:
                return tuple(v for v in (embeds, last, hidden) if v is not None)
            return _clip_output("CLIPTextModelOutput", text_embeds=embeds, last_hidden_state=last,
                                hidden_states=hidden)
        if not return_dict:
            return tuple(v for v in (last, pooled, hidden) if v is not None)
        return _clip_output("BaseModelOutputWithPooling", last_hidden_state=last, pooler_output=pooled,
                            hidden_states=hidden)


class CompiledVisionEncoder:
    """transformers ``CLIPVisionModel`` / ``CLIPVisionModelWithProjection`` forward on the native path (the
    SVD pipeline's image_encoder, reference :100-103).  One plan + CUDA graph per batch size."""

    def __init__(self, config, state_dict_fn, eager_forward, with_projection, enable_cuda_graph=True):
        from .clip_plan import clip_vision_spec_from_config
        self.spec = clip_vision_spec_from_config(config, with_projection)
        self._state_dict_fn, self._eager = state_dict_fn, eager_forward
        self.with_projection = with_projection
        self.enable_cuda_graph = enable_cuda_graph
        self._weights, self._param_refs, self._param_versions = None, None, None
        self._cached = {}
        self._lock = threading.Lock()
        self._warned = False
        rd = cfg_get(config, "return_dict", None)
        self._return_dict_default = True if rd is None else bool(rd)

    def _route_eager(self, why, call):
        if not self._warned:
            logger.warning("sfast (B200 build): image encoder call left on the module's eager path (%s)", why)
            self._warned = True
        return self._eager(**call)

    def __call__(self, pixel_values=None, interpolate_pos_encoding=False, output_attentions=None,
                 output_hidden_states=None, return_dict=None, **kwargs):
        from .clip_plan import ClipVisionPlan
        call = dict(pixel_values=pixel_values, interpolate_pos_encoding=interpolate_pos_encoding,
                    output_attentions=output_attentions, output_hidden_states=output_hidden_states,
                    return_dict=return_dict, **kwargs)
        call = {k: v for k, v in call.items() if v is not None}
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        spec = self.spec
        if interpolate_pos_encoding or output_attentions or kwargs:
            return self._route_eager("interpolate_pos_encoding / output_attentions / extra kwargs", call)
        if tuple(pixel_values.shape[1:]) != (spec.channels, spec.image_size, spec.image_size):
            raise ValueError(f"Input image size {tuple(pixel_values.shape[2:])} doesn't match model "
                             f"({spec.image_size}*{spec.image_size}).")
        sd = None
        if self._weights is None:
            sd = self._state_dict_fn()
            wdt = sd["vision_model.embeddings.patch_embedding.weight"].dtype
            if wdt not in (torch.float16, torch.bfloat16):
                return self._route_eager(f"{wdt} weights: the native encoder computes in fp16 / bf16", call)
        require_b200(pixel_values.device)
        B = pixel_values.shape[0]
        with self._lock, torch.cuda.device(pixel_values.device):
            if self._weights is None:
                wdt = sd["vision_model.embeddings.patch_embedding.weight"].dtype
                self._weights = PackedWeights(self.spec, sd, wdt, pixel_values.device)
                self._param_refs = [t for t in sd.values() if torch.is_tensor(t)]
                self._param_versions = [t._version for t in self._param_refs]
            elif [t._version for t in self._param_refs] != self._param_versions:
                sd = self._state_dict_fn()
                torch.cuda.current_stream().synchronize()
                self._weights.refresh(sd)
                self._param_refs = [t for t in sd.values() if torch.is_tensor(t)]
                self._param_versions = [t._version for t in self._param_refs]
            key = (B, pixel_values.device.index)
            gp = self._cached.get(key)
            if gp is None:
                gp = _GraphedPlan(ClipVisionPlan(self._weights, B), self.enable_cuda_graph)
                self._cached[key] = gp
            plan = gp.plan
            plan.pixels_in.copy_(pixel_values, non_blocking=True)   # (casts to the module's dtype, as the module does)
            gp.step()
            last = plan.last_hidden_state.clone()
            pooled = plan.pooled.clone()
            embeds = plan.image_embeds.clone() if self.with_projection else None
            hidden = tuple(h.clone() for h in plan.hidden_states()) if output_hidden_states else None
        if return_dict is None:
            return_dict = self._return_dict_default
        if self.with_projection:
            if not return_dict:
                return tuple(v for v in (embeds, last, hidden) if v is not None)
            return _clip_output("CLIPVisionModelOutput", image_embeds=embeds, last_hidden_state=last,
                                hidden_states=hidden)
        if not return_dict:
            return tuple(v for v in (last, pooled, hidden) if v is not None)
        return _clip_output("BaseModelOutputWithPooling", last_hidden_state=last, pooler_output=pooled,
                            hidden_states=hidden)


def compile_image_encoder_module(m, enable_cuda_graph=True):
    """Replace ``m.forward`` of a CLIPVisionModel / CLIPVisionModelWithProjection (same module object)."""
    with_projection = hasattr(m, "visual_projection")
    eager = m.forward
    compiled = CompiledVisionEncoder(m.config, m.state_dict, eager, with_projection, enable_cuda_graph)

    def forward(*args, **kwargs):
        if args:
            kwargs = dict(zip(("pixel_values",), args), **kwargs)
        return compiled(**kwargs)

    forward.__self__ = m
    forward._cached = compiled._cached
    forward._compiled = compiled
    m.forward = forward
    return m


class _PlainOutput(dict):
    """Stand-in for transformers' ModelOutput when transformers is not importable: attribute, key and
    integer access over the non-None fields, like the real class."""

    def __init__(self, **kw):
        super().__init__({k: v for k, v in kw.items() if v is not None})
        self.__dict__.update(kw)

    def __getitem__(self, k):
        return tuple(self.values())[k] if isinstance(k, (int, slice)) else super().__getitem__(k)


def _clip_output(kind, **fields):
    try:
        if kind == "CLIPTextModelOutput":
            from transformers.models.clip.modeling_clip import CLIPTextModelOutput as cls
        elif kind == "CLIPVisionModelOutput":
            from transformers.models.clip.modeling_clip import CLIPVisionModelOutput as cls
        else:
            from transformers.modeling_outputs import BaseModelOutputWithPooling as cls
        return cls(**fields)
    except Exception:  # noqa: BLE001
        return _PlainOutput(**fields)


def compile_text_encoder_module(m, enable_cuda_graph=True):
    """Replace ``m.forward`` of a CLIPTextModel / CLIPTextModelWithProjection (same module object)."""
    with_projection = hasattr(m, "text_projection")
    eager = m.forward
    compiled = CompiledTextEncoder(m.config, m.state_dict, eager, with_projection, enable_cuda_graph)

    def forward(*args, **kwargs):
        if args:
            kwargs = dict(zip(("input_ids", "attention_mask", "position_ids"), args), **kwargs)
        return compiled(**kwargs)

    forward.__self__ = m
    forward._cached = compiled._cached
    forward._compiled = compiled
    m.forward = forward
    return m


def compile_unet_module(m, enable_cuda_graph=True, preserve_parameters=True):
    """Replace ``m.forward`` (same module object, as the reference does at
    /root/reference/src/sfast/compilers/diffusion_pipeline_compiler.py:146-149)."""
    compiled = CompiledUNet(m.config, m.state_dict, enable_cuda_graph, preserve_parameters)

    def forward(*args, **kwargs):
        return compiled(*args, **kwargs)

    forward.__self__ = m
    forward._cached = compiled._cached
    forward._compiled = compiled
    m.forward = forward
    return m
