"""Drop-in `Dreamer` module: the reference's API (pydreamer/models/dreamer.py:19-230) over hand-written
sm_100a kernels.

What is kept from the reference contract:
  * ctor `Dreamer(conf)` with the reference's config keys; an nn.Module whose state_dict keys and
    shapes equal the reference's, so checkpoints round-trip with an unmodified reference Dreamer;
  * `init_state`, `training_step` (same arguments, same 5-tuple, same dict keys), `init_optimizers`
    (same tuple order), `grad_clip` (same dict keys); each returned loss is backwarded separately by
    the caller exactly as train.py:184-187 does.

What is different inside: the whole step — forward AND backward — is an explicit schedule of kernels
from libpd_b200.so (no autograd tape, no torch math).  `training_step` computes the gradients of the
four losses directly into a flat fp32 gradient arena; the returned loss tensors are connected to the
parameters through a tiny autograd.Function whose backward hands those precomputed gradients over,
so `loss.backward()` in the caller works unchanged.  Parameters live in one flat arena (views), which
is what the fused optimizer and the single-bucket data-parallel all-reduce operate on.
"""
import math
import contextlib
import os
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops as _ops
from .ops import ACT_ELU, ACT_NONE


# ======================================================================================
# parameter containers (names/shapes identical to the reference's module tree)
# ======================================================================================
def _mlp_seq(in_dim, out_dim, hidden_dim, hidden_layers, layer_norm):
    """common.py:37-65"""
    if not layer_norm:
        raise NotImplementedError("layer_norm=False is outside the accelerated path (SURVEY.md §8f N4)")
    layers, dim = [], in_dim
    for _ in range(hidden_layers):
        layers += [nn.Linear(dim, hidden_dim), nn.LayerNorm(hidden_dim, eps=1e-3), nn.ELU()]
        dim = hidden_dim
    layers += [nn.Linear(dim, out_dim)]
    if out_dim == 1:
        layers += [nn.Flatten(0)]
    return nn.Sequential(*layers)


class _MLP(nn.Module):
    def __init__(self, in_dim, out_dim, hidden_dim, hidden_layers, layer_norm):
        super().__init__()
        self.in_dim, self.out_dim, self.hidden_dim, self.hidden_layers = in_dim, out_dim, hidden_dim, hidden_layers
        self.model = _mlp_seq(in_dim, out_dim, hidden_dim, hidden_layers, layer_norm)


class _DenseHead(nn.Module):
    """decoders.py:257-319 (DenseBernoulliDecoder / DenseNormalDecoder): `.model` is an MLP"""

    def __init__(self, in_dim, hidden_layers, layer_norm, hidden_dim=400):
        super().__init__()
        self.model = _MLP(in_dim, 1, hidden_dim, hidden_layers, layer_norm)


class _ConvEncoder(nn.Module):
    """encoders.py:72-96"""

    def __init__(self, in_channels, d):
        super().__init__()
        self.out_dim = d * 32
        self.model = nn.Sequential(nn.Conv2d(in_channels, d, 4, 2), nn.ELU(), nn.Conv2d(d, d * 2, 4, 2), nn.ELU(),
                                   nn.Conv2d(d * 2, d * 4, 4, 2), nn.ELU(), nn.Conv2d(d * 4, d * 8, 4, 2), nn.ELU(),
                                   nn.Flatten())


class _MultiEncoder(nn.Module):
    def __init__(self, conf):
        super().__init__()
        if conf.image_encoder != "cnn" or conf.reward_input or conf.vecobs_size:
            raise NotImplementedError("accelerated path covers image_encoder=cnn without reward_input/vecobs "
                                      "(SURVEY.md §2 row 4); use the reference for other encoders")
        self.encoder_image = _ConvEncoder(conf.image_channels, conf.cnn_depth)
        self.out_dim = self.encoder_image.out_dim


class _ConvDecoder(nn.Module):
    """decoders.py:111-161"""

    def __init__(self, in_dim, out_channels, d):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(in_dim, d * 32), nn.Unflatten(-1, (d * 32, 1, 1)),
                                   nn.ConvTranspose2d(d * 32, d * 4, 5, 2), nn.ELU(),
                                   nn.ConvTranspose2d(d * 4, d * 2, 5, 2), nn.ELU(),
                                   nn.ConvTranspose2d(d * 2, d, 6, 2), nn.ELU(),
                                   nn.ConvTranspose2d(d, out_channels, 6, 2))


class _MultiDecoder(nn.Module):
    def __init__(self, features_dim, conf):
        super().__init__()
        if conf.image_decoder != "cnn" or conf.reward_decoder_categorical or conf.vecobs_size:
            raise NotImplementedError("accelerated path covers image_decoder=cnn + Normal reward head "
                                      "(SURVEY.md §2 row 5)")
        if conf.image_size != 64:
            raise NotImplementedError("conv geometry is the reference's 64x64 one (encoders.py:77-90)")
        self.image = _ConvDecoder(features_dim, conf.image_channels, conf.cnn_depth)
        self.reward = _DenseHead(features_dim, conf.reward_decoder_layers, conf.layer_norm)
        self.terminal = _DenseHead(features_dim, conf.terminal_decoder_layers, conf.layer_norm)


class _GRUStack(nn.Module):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.layers = nn.ModuleList([nn.GRUCell(input_size, hidden_size)])


class _RSSMCell(nn.Module):
    """rssm.py:96-116"""

    def __init__(self, embed_dim, action_dim, deter_dim, stoch_dim, stoch_discrete, hidden_dim):
        super().__init__()
        z = stoch_dim * stoch_discrete
        self.z_mlp = nn.Linear(z, hidden_dim)
        self.a_mlp = nn.Linear(action_dim, hidden_dim, bias=False)
        self.in_norm = nn.LayerNorm(hidden_dim, eps=1e-3)
        self.gru = _GRUStack(hidden_dim, deter_dim)
        self.prior_mlp_h = nn.Linear(deter_dim, hidden_dim)
        self.prior_norm = nn.LayerNorm(hidden_dim, eps=1e-3)
        self.prior_mlp = nn.Linear(hidden_dim, z)
        self.post_mlp_h = nn.Linear(deter_dim, hidden_dim)
        self.post_mlp_e = nn.Linear(embed_dim, hidden_dim, bias=False)
        self.post_norm = nn.LayerNorm(hidden_dim, eps=1e-3)
        self.post_mlp = nn.Linear(hidden_dim, z)


class _RSSMCore(nn.Module):
    def __init__(self, *a):
        super().__init__()
        self.cell = _RSSMCell(*a)


def _init_weights_tf2(m):
    """functions.py:81-94"""
    if type(m) in (nn.Conv2d, nn.ConvTranspose2d, nn.Linear):
        nn.init.xavier_uniform_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)
    if type(m) == nn.GRUCell:
        nn.init.xavier_uniform_(m.weight_ih.data)
        nn.init.orthogonal_(m.weight_hh.data)
        nn.init.zeros_(m.bias_ih.data)
        nn.init.zeros_(m.bias_hh.data)


class _WorldModel(nn.Module):
    """dreamer.py:232-284"""

    def __init__(self, conf):
        super().__init__()
        if not conf.stoch_discrete:
            raise NotImplementedError("Gaussian latents (stoch_discrete=0) are outside the accelerated path (§8f N4)")
        if conf.gru_layers != 1 or conf.gru_type != "gru":
            raise NotImplementedError("accelerated path covers gru_type=gru, gru_layers=1 (SURVEY.md §2 row 3)")
        if conf.aux_critic:
            raise NotImplementedError("aux_critic is outside the accelerated path")
        if conf.stoch_discrete > 32 or conf.stoch_dim > 32:
            raise NotImplementedError("categorical kernels handle <= 32 groups of <= 32 classes")
        self.encoder = _MultiEncoder(conf)
        features_dim = conf.deter_dim + conf.stoch_dim * conf.stoch_discrete
        self.decoder = _MultiDecoder(features_dim, conf)
        self.core = _RSSMCore(self.encoder.out_dim, conf.action_dim, conf.deter_dim, conf.stoch_dim,
                              conf.stoch_discrete, conf.hidden_dim)
        for m in self.modules():
            _init_weights_tf2(m)


class _ActorCritic(nn.Module):
    """a2c.py:11-41"""

    def __init__(self, in_dim, out_actions, layer_norm, actor_dist, hidden_dim=400, hidden_layers=4):
        super().__init__()
        actor_out = out_actions if actor_dist == "onehot" else 2 * out_actions
        self.actor = _MLP(in_dim, actor_out, hidden_dim, hidden_layers, layer_norm)
        self.critic = _MLP(in_dim, 1, hidden_dim, hidden_layers, layer_norm)
        self.critic_target = _MLP(in_dim, 1, hidden_dim, hidden_layers, layer_norm)
        self.critic_target.requires_grad_(False)
        self.train_steps = 0


class _NoProbeHead(nn.Module):
    """probes.py:140-150"""

    def __init__(self):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1), requires_grad=True)


# ======================================================================================
# loss <-> precomputed-gradient bridge
# ======================================================================================
class _AttachGrads(torch.autograd.Function):
    """Returns `value` as a differentiable scalar whose backward delivers the gradients that the kernel
    schedule already wrote into the arena group `gid` (scaled by the incoming grad_output)."""

    @staticmethod
    def forward(ctx, value, owner, gid, *params):
        ctx.owner, ctx.gid = owner, gid
        return value.clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.owner._deliver_grads(ctx.gid, grad_out)
        return (None, None, None) + tuple(None for _ in ctx.owner._group_params[ctx.gid])


class _FusedAdamW:
    """One optimizer of the tuple `init_optimizers` returns (dreamer.py:60-71): torch.optim.AdamW
    semantics (decoupled weight_decay=0.01, eps, no amsgrad) over one contiguous arena group.

    `state_dict()` / `load_state_dict()` speak torch.optim.AdamW's own layout (per-parameter `state` keyed by the
    parameter's index in the group, `param_groups[0]['params']` = those indices), so the reference's checkpoint code
    (tools.py:171-172 save, :195-196 load) round-trips both ways with a `torch.optim.AdamW` over the reference module."""

    def __init__(self, owner, gid, lr, eps, betas=(0.9, 0.999), weight_decay=0.01):
        self.owner, self.gid = owner, gid
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                             foreach=None, capturable=False, differentiable=False, fused=None)
        self.param_groups = [dict(params=list(owner._group_params[gid]), **self.defaults)]
        self._alloc()

    def _alloc(self):
        a = self.owner._group_slice(self.gid, self.owner._arena)
        self.exp_avg = torch.zeros_like(a)
        self.exp_avg_sq = torch.zeros_like(a)
        self.step_t = torch.zeros(1, dtype=torch.int32, device=a.device)

    def zero_grad(self, set_to_none=True):
        # every training_step overwrites the gradient arena, so there is nothing to clear; the call marks the previous
        # gradients as consumed (see Dreamer._note_new_grads: accumulation over several backward passes is not supported)
        self.owner._grads_pending.discard(self.gid)
        return None

    @torch.no_grad()
    def step(self, closure=None):
        o = self.owner
        o._ensure_arena()
        p = o._group_slice(self.gid, o._arena)
        g = o._group_slice(self.gid, o._garena)
        if self.exp_avg.device != p.device:
            self._alloc()
        pg = self.param_groups[0]
        o.ops.inc(self.step_t)
        o.ops.adamw(p, g, self.exp_avg, self.exp_avg_sq, pg["lr"], pg["betas"][0], pg["betas"][1], pg["eps"],
                    pg["weight_decay"], self.step_t)
        o._weights_dirty = True
        o._grads_pending.discard(self.gid)

    def _slices(self):
        """(index, offset in the group, numel, shape) of every parameter of the group, in torch's parameter order."""
        o = self.owner
        base = o._group_range[self.gid][0]
        return [(i, o._offsets[id(p)] - base, p.numel(), p.shape) for i, p in enumerate(o._group_params[self.gid])]

    def state_dict(self):
        step = int(self.step_t.item())
        state = {}
        if step > 0:                                          # torch creates the per-parameter state at the first step
            for i, off, n, shape in self._slices():
                state[i] = dict(step=torch.tensor(float(step)), exp_avg=self.exp_avg[off:off + n].view(shape).clone(),
                                exp_avg_sq=self.exp_avg_sq[off:off + n].view(shape).clone())
        pg = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        pg["params"] = list(range(len(self.owner._group_params[self.gid])))
        return dict(state=state, param_groups=[pg])

    def load_state_dict(self, sd):
        st = sd["state"]
        sl = self._slices()
        if len(sd["param_groups"]) != 1 or len(sd["param_groups"][0]["params"]) != len(sl):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        ids = list(sd["param_groups"][0]["params"])
        with torch.no_grad():
            self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self.step_t.zero_()
            steps = set()
            for (i, off, n, shape), key in zip(sl, ids):
                e = st.get(key)
                if e is None:
                    continue
                if tuple(e["exp_avg"].shape) != tuple(shape):
                    raise ValueError(f"optimizer state of parameter {i} has shape {tuple(e['exp_avg'].shape)}, expected {tuple(shape)}")
                self.exp_avg[off:off + n].copy_(e["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(e["exp_avg_sq"].reshape(-1))
                steps.add(int(float(e["step"])))
            if len(steps) > 1:
                raise ValueError(f"parameters of one group carry different step counts {sorted(steps)}: the fused optimizer "
                                 "keeps one counter per group")
            if steps:
                self.step_t.fill_(steps.pop())
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v


# ======================================================================================
# the module
# ======================================================================================
GROUPS = ("wm", "probe", "actor", "critic")


class Dreamer(nn.Module):

    def __init__(self, conf):
        super().__init__()
        assert conf.action_dim > 0, "Need to set action_dim to match environment"   # dreamer.py:23
        if conf.probe_model != "none":
            raise NotImplementedError("probe heads are research probes outside the hot path (SURVEY.md §2 row 9)")
        if conf.probe_gradients:
            raise NotImplementedError("probe_gradients is for the baselines (SURVEY.md §2 row 10)")
        if conf.actor_dist not in ("onehot", "tanh_normal"):
            raise NotImplementedError(f"actor_dist={conf.actor_dist}")
        if conf.actor_grad != "reinforce":
            raise NotImplementedError("actor_grad=dynamics asserts upstream at a2c.py:131 (SURVEY.md §0.5); "
                                      "the accelerated path implements reinforce")
        self.conf = conf
        self.iwae_samples = conf.iwae_samples
        self.imag_horizon = conf.imag_horizon
        self.probe_gradients = conf.probe_gradients
        features_dim = conf.deter_dim + conf.stoch_dim * conf.stoch_discrete
        self.wm = _WorldModel(conf)
        self.ac = _ActorCritic(features_dim, conf.action_dim, conf.layer_norm, conf.actor_dist)
        self.probe_model = _NoProbeHead()
        # static dims
        self.d = SimpleNamespace(D=conf.deter_dim, G=conf.stoch_dim, C=conf.stoch_discrete,
                                 Z=conf.stoch_dim * conf.stoch_discrete, Hd=conf.hidden_dim, E=self.wm.encoder.out_dim,
                                 A=conf.action_dim, F=features_dim, cd=conf.cnn_depth, IC=conf.image_channels,
                                 Aout=conf.action_dim if conf.actor_dist == "onehot" else 2 * conf.action_dim)
        self._arena = None
        self._arena_device = None
        self._ws = {}
        self._graphs = {}
        self._ops = None
        self._weights_dirty = True
        self._dp = None           # optional data-parallel reducer (pydreamer_b200.parallel)
        self._grads_pending = set()
        # nn.Module.load_state_dict copies into the arena views in place: the tf32 / fp16 shadow arenas and the re-laid
        # conv weights must be rebuilt before the next kernel reads them
        self.register_load_state_dict_post_hook(lambda module, incompatible: setattr(module, "_weights_dirty", True))
        self._build_registry()

    # ------------------------------------------------------------------ arena
    def _build_registry(self):
        self._group_params = {
            "wm": list(self.wm.parameters()),
            "probe": list(self.probe_model.parameters()),
            "actor": list(self.ac.actor.parameters()),
            "critic": list(self.ac.critic.parameters()),
            "target": list(self.ac.critic_target.parameters()),
        }
        self._names = {id(p): n for n, p in self.named_parameters()}
        off = 0
        self._offsets, self._group_range = {}, {}
        for gname in GROUPS + ("target",):
            start = off
            for p in self._group_params[gname]:
                self._offsets[id(p)] = off
                off += (p.numel() + 7) // 8 * 8          # every tensor 16-byte aligned in the fp32 AND the fp16 arena (TMA)
            self._group_range[gname] = (start, off)
        self._arena_numel = off
        self._train_numel = self._group_range["critic"][1]

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._arena = None   # .to()/.cuda() re-created the tensors: re-flatten lazily
        return r

    def _ensure_arena(self):
        p0 = next(self.parameters())
        if self._arena is not None and self._arena_device == p0.device and p0.data_ptr() == self._arena.data_ptr():
            return
        dev = p0.device
        arena = torch.zeros(self._arena_numel, dtype=torch.float32, device=dev)
        garena = torch.zeros(self._train_numel, dtype=torch.float32, device=dev)
        for g in GROUPS + ("target",):
            for p in self._group_params[g]:
                o = self._offsets[id(p)]
                v = arena[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                if g != "target":
                    p.grad = garena[o:o + p.numel()].view(p.shape)
        self._arena, self._garena, self._arena_device = arena, garena, dev
        self._sarena = torch.zeros_like(arena)     # tf32-rounded shadow of the arena (GEMM operands)
        self._harena = torch.zeros(arena.shape, dtype=torch.float16, device=dev)   # fp16 shadow (forward-only GEMMs)
        self._ws = {}
        self._graphs = {}
        self._ops = None
        self._weights_dirty = True

    def _group_slice(self, gid, arena):
        a, b = self._group_range[gid]
        return arena[a:b]

    @property
    def ops(self):
        if self._ops is None:
            self._ops = _ops.get_ops(next(self.parameters()).device)
        return self._ops

    def _view(self, arena, p):
        o = self._offsets[id(p)]
        return arena[o:o + p.numel()].view(p.shape)

    def _deliver_grads(self, gid, grad_out):
        """backward() of one of the four losses: scale the group's precomputed gradients by grad_out (a device-side
        no-op when it is exactly 1, the plain `loss.backward()` of train.py:186-187; a GradScaler under amp or any
        `(k * loss).backward()` scales them) and make sure `.grad` of its parameters points at them."""
        g = self._group_slice(gid, self._garena)
        self.ops.scale_by(g, grad_out.reshape(1).to(device=g.device, dtype=g.dtype))
        for p in self._group_params[gid]:
            if p.grad is None or p.grad.data_ptr() != self._view(self._garena, p).data_ptr():
                p.grad = self._view(self._garena, p)

    def _note_new_grads(self):
        """Every training_step overwrites the gradient arena (it is never accumulated into).  A caller that runs two
        training steps without an optimizer step / zero_grad in between is accumulating gradients in the reference's
        semantics; that is not supported here and is said once instead of silently training on the last micro-batch."""
        if self._grads_pending and not self._warned_accum:
            import warnings
            warnings.warn("pydreamer_b200: training_step overwrites the gradients of the previous call (groups "
                          f"{sorted(self._grads_pending)} were neither stepped nor zero_grad()-ed): gradient accumulation "
                          "over several training_step calls is not supported")
            self._warned_accum = True
        self._grads_pending = set(GROUPS)

    _warned_accum = False
    # forward-only layers (imagination rollout, heads on dreamed features) use fp16 tensor-core operands: the same
    # 10-bit mantissa as TF32 at twice the MMA rate and half the operand traffic; no gradient flows through them.
    fp16_forward = os.environ.get("PD_B200_FP16_FORWARD", "1") != "0"
    # The decoder's deconvolution column matrices (GEMM output -> col2im fold; written once, read once, ~4 GB per step in fp32)
    # are stored in fp16 on the fp16-forward product path: half the HBM traffic of the two kernels either side of them.
    fp16_cols = os.environ.get("PD_B200_FP16_COLS", "0") != "0"
    # conv / deconv contractions gather their operand with TMA im2col-mode loads (pd_conv_gemm) instead of materialising
    # im2col matrices: encoder layers 2-4 (forward + weight gradient), deconv layers 2-3 (input + weight gradient).
    implicit_conv = os.environ.get("PD_B200_IMPLICIT_CONV", "1") != "0"

    # ------------------------------------------------------------------ reference API
    def init_optimizers(self, lr, lr_actor=None, lr_critic=None, eps=1e-5):
        self._ensure_arena()
        return (_FusedAdamW(self, "wm", lr, eps), _FusedAdamW(self, "probe", lr, eps),
                _FusedAdamW(self, "actor", lr_actor or lr, eps), _FusedAdamW(self, "critic", lr_critic or lr, eps))

    @torch.no_grad()
    def grad_clip(self, grad_clip, grad_clip_ac=None):
        """dreamer.py:73-87: per-group clip_grad_norm_, returns the pre-clip norms."""
        self._ensure_arena()
        if self._dp is not None:
            self._dp.allreduce_grads(self)
        ws = self._buf("clip", 8)
        self.ops.fill(ws, 0.0)
        keys = []
        for i, (gid, key, mx) in enumerate((("wm", "grad_norm", grad_clip), ("probe", "grad_norm_probe", grad_clip),
                                            ("actor", "grad_norm_actor", grad_clip_ac or grad_clip),
                                            ("critic", "grad_norm_critic", grad_clip_ac or grad_clip))):
            g = self._group_slice(gid, self._garena)
            self.ops.sumsq(g, ws[i:i + 1])
            self.ops.clip_scale(g, ws[i:i + 1], mx, ws[4 + i:5 + i])
            keys.append(key)
        norms = ws[4:8].clone()                 # the caller's own copy (the workspace is rewritten by the next call)
        return {k: norms[i] for i, k in enumerate(keys)}

    def init_state(self, batch_size):
        dev = next(self.parameters()).device
        return (torch.zeros((batch_size, self.d.D), device=dev), torch.zeros((batch_size, self.d.Z), device=dev))

    @torch.no_grad()
    def inference(self, obs, in_state):
        """dreamer.py:92-111: one posterior step (T=1) + actor / critic forward for the env-interaction policy.
        Returns (action distribution, out_state, {'policy_value'}) like the reference (generator.py:321-328)."""
        import torch.distributions as D

        assert "action" in obs, "Observation should contain previous action"
        act_shape = obs["action"].shape
        assert len(act_shape) == 3 and act_shape[0] == 1, f"Expected shape (1,B,A), got {act_shape}"
        self._ensure_arena()
        self._prepare_weights()
        d, B = self.d, act_shape[1]
        noise = self._buf("inf.noise", 1, B, d.Z).exponential_()
        if getattr(self, "_test_inference_noise", None) is not None:      # tests pin the sampling noise
            noise = self._test_inference_noise
        feat = self._buf("inf.feat", 1, B, d.F)
        _, _, _, out_state = self._wm_features(obs, in_state, 1, B, 1, noise, "inf.", feat)
        f = feat.view(B, d.F)
        alog, val = self._buf("inf.alog", B, d.Aout), self._buf("inf.val", B, 1)
        self._mlp_fwd(self._mlp_params(self.ac.actor), f, alog, "scratch")
        self._mlp_fwd(self._mlp_params(self.ac.critic), f, val, "scratch")
        y = alog.view(1, B, d.Aout).clone()
        if self.conf.actor_dist == "onehot":
            dist = D.OneHotCategorical(logits=y)
        else:                                                          # functions.py:69-78
            mean = 5 * torch.tanh(y[..., :d.A] / 5)
            std = torch.nn.functional.softplus(y[..., d.A:]) + 0.1
            normal = D.independent.Independent(D.normal.Normal(mean, std), 1)
            dist = D.TransformedDistribution(normal, [D.TanhTransform()])
            dist.entropy = normal.entropy
        return dist, out_state, dict(policy_value=val.mean())

    # ------------------------------------------------------------------ workspace
    def _buf(self, name, *shape, dtype=torch.float32, zero=False):
        key = (name, shape, dtype)
        t = self._ws.get(key)
        if t is None:
            dev = self._arena.device
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=dev)
            self._ws[key] = t
        return t

    # ------------------------------------------------------------------ weights
    def _w(self, p):      # rounded shadow (tensor-core operand)
        return self._view(self._sarena, p)

    def _wh(self, p):     # fp16 shadow (operand of forward-only GEMMs)
        return self._view(self._harena, p)

    def _raw(self, p):    # fp32 master (biases, LayerNorm affine)
        return self._view(self._arena, p)

    def _g(self, p):
        return self._view(self._garena, p)

    def _w2d(self, p):
        return self._w(p).view(p.shape[0], -1)

    def _prepare_weights(self):
        """Per optimizer step: tf32-round the whole arena into the shadow arena (one kernel) and build the
        conv / deconv weights in GEMM layout: conv (Cout,Cin,kh,kw)->(Cout,(kh,kw,Cin)); deconv
        (Cin,Cout,kh,kw)->((kh,kw,Cout),Cin)."""
        if not self._weights_dirty:
            return
        ops = self.ops
        ops.round_copy(self._arena, self._sarena, True)
        if self.fp16_forward:
            ops.to_half(self._arena.view(1, -1), self._harena.view(1, -1))
        enc = self.wm.encoder.encoder_image.model
        self._encw = []
        for li, idx in enumerate((0, 2, 4, 6)):
            w = enc[idx].weight
            if li == 0:
                self._encw.append(self._w(w).view(w.shape[0], -1))          # (c,kh,kw) order == native
            else:
                co, ci, kh, kw = w.shape
                sh = self._buf(f"encw{li}", co, kh, kw, ci)
                ops.permute4(self._w(w), sh, (0, 2, 3, 1))
                self._encw.append(sh.view(co, kh * kw * ci))
        dec = self.wm.decoder.image.model
        self._decw = []
        for li, idx in enumerate((2, 4, 6, 8)):
            w = dec[idx].weight
            ci, co, kh, kw = w.shape
            sh = self._buf(f"decw{li}", kh, kw, co, ci)
            ops.permute4(self._w(w), sh, (2, 3, 1, 0))
            self._decw.append(sh.view(kh * kw * co, ci))
        if self.persistent_rssm:      # z_mlp^T [Z, Hd] fp16: a one-hot latent selects rows (persistent unroll, phase A)
            wz = self.wm.core.cell.z_mlp.weight
            self._k1_wzT = self._buf("k1.wzT", wz.shape[1], wz.shape[0], dtype=torch.float16)
            ops.transpose_to_half(self._raw(wz), self._k1_wzT)
        if self.persistent_bptt:      # transposed fp16 copies: operands of the persistent BPTT kernel (pd_rssm_unroll_bwd)
            cell = self.wm.core.cell
            gru = cell.gru.layers[0]
            self._k1b_w = {}
            for name, wgt in (("w_pmT16", cell.post_mlp.weight), ("w_phT16", cell.post_mlp_h.weight),
                              ("w_hhT16", gru.weight_hh), ("w_ihT16", gru.weight_ih), ("w_zT16", cell.z_mlp.weight)):
                buf = self._buf("k1b." + name, wgt.shape[1], wgt.shape[0], dtype=torch.float16)
                ops.transpose_to_half(self._raw(wgt), buf)
                self._k1b_w[name] = buf
        # a_mlp^T [A, Hd]: a one-hot action selects one row (imagination rollout, pd_gather_rows)
        wa = self.wm.core.cell.a_mlp.weight
        self._waT = self._buf("waT", wa.shape[1], wa.shape[0])
        ops.permute4(self._w(wa).view(wa.shape[0], wa.shape[1], 1, 1), self._waT.view(wa.shape[1], wa.shape[0], 1, 1),
                     (1, 0, 2, 3))
        self._weights_dirty = False

    def _mlp_params(self, mlp):
        seq = mlp.model
        L = mlp.hidden_layers
        return SimpleNamespace(L=L, lin=[seq[3 * l] for l in range(L)], ln=[seq[3 * l + 1] for l in range(L)],
                               out=seq[3 * L], hid=mlp.hidden_dim, out_dim=mlp.out_dim, in_dim=mlp.in_dim)

    # ------------------------------------------------------------------ MLP forward / backward
    def _mlp_fwd(self, mp, x_in, out, tag, rows_total=None, row0=0, save=False, x16=None):
        """out[rows, out_dim] = MLP(x_in).  With save=True the per-layer pre-norm x, post-ELU y and LN
        statistics are kept in workspace buffers `tag` (rows_total rows, this call fills [row0, row0+rows)).
        x16 (optional, fp16 copy of x_in): run the hidden-layer GEMMs with fp16 operands (forward-only use)."""
        ops = self.ops
        rows = x_in.shape[0]
        RT = rows_total or rows
        inp, inp16 = x_in, x16
        f16 = x16 is not None
        for l in range(mp.L):
            if save:
                x = self._buf(f"{tag}.x{l}", RT, mp.hid)[row0:row0 + rows]
                y = self._buf(f"{tag}.y{l}", RT, mp.hid)[row0:row0 + rows]
                mean = self._buf(f"{tag}.m{l}", RT)[row0:row0 + rows]
                rstd = self._buf(f"{tag}.r{l}", RT)[row0:row0 + rows]
            else:
                x = self._buf(f"{self._scratch_ns}mlp.sx", rows, mp.hid)
                y = self._buf(f"{self._scratch_ns}mlp.sy{l % 2}", rows, mp.hid)
                mean = self._buf(f"{self._scratch_ns}mlp.sm", rows)
                rstd = self._buf(f"{self._scratch_ns}mlp.sr", rows)
            if f16:
                ops.gemm_f16(inp16, self._wh(mp.lin[l].weight), x, bias=self._raw(mp.lin[l].bias))
                y16 = self._buf(f"{self._scratch_ns}mlp.h16_{l % 2}", rows, mp.hid, dtype=torch.float16)
            else:
                ops.gemm(inp, self._w(mp.lin[l].weight), x, bias=self._raw(mp.lin[l].bias))
                y16 = None
            ops.ln_elu_fwd(x, self._raw(mp.ln[l].weight), self._raw(mp.ln[l].bias), 1e-3, y, mean, rstd, y16)
            inp, inp16 = y, y16
        ops.gemm(inp, self._w(mp.out.weight), out, bias=self._raw(mp.out.bias))     # narrow output layer: fp32 operands
        return out

    def _mlp_bwd(self, mp, x_in, dout, tag, rows_total=None, din=None, din_accum=False):
        """Accumulates parameter grads of the MLP; optionally (+)= the input gradient into din."""
        ops = self.ops
        rows = x_in.shape[0]
        RT = rows_total or rows
        sv = lambda n, l: self._buf(f"{tag}.{n}{l}", RT, mp.hid)[:rows]
        st = lambda n, l: self._buf(f"{tag}.{n}{l}", RT)[:rows]
        dy = self._buf(f"{self._scratch_ns}mlp.dy", rows, mp.hid)
        dx = self._buf(f"{self._scratch_ns}mlp.dx", rows, mp.hid)
        ylast = sv("y", mp.L - 1)
        ops.gemm(dout, ylast, self._g(mp.out.weight), a_mn=True, b_mn=True, accumulate=True)
        ops.colsum(dout, self._g(mp.out.bias))
        ops.gemm(dout, self._w(mp.out.weight), dy, b_mn=True)
        for l in reversed(range(mp.L)):
            ops.ln_elu_bwd(dy, sv("x", l), sv("y", l), self._raw(mp.ln[l].weight), st("m", l), st("r", l), dx,
                           self._g(mp.ln[l].weight), self._g(mp.ln[l].bias), self._g(mp.lin[l].bias))
            inp = x_in if l == 0 else sv("y", l - 1)
            ops.gemm(dx, inp, self._g(mp.lin[l].weight), a_mn=True, b_mn=True, accumulate=True)
            if l > 0:
                ops.gemm(dx, self._w(mp.lin[l].weight), dy, b_mn=True)
            elif din is not None:
                ops.gemm(dx, self._w(mp.lin[0].weight), din, b_mn=True, res=din if din_accum else None)

    # ------------------------------------------------------------------ noise
    def _draw_noise(self, T, BI, N, H, image_pred=False, dream_log=False, B=0):
        """Exp(1) noise for the categorical samples in the reference's consumption order;
        Gaussian noise for the tanh_normal actor."""
        d, dev = self.d, self._arena.device
        post = self._buf("noise.post", T, BI, d.Z).exponential_()
        if self.conf.actor_dist == "onehot":
            actor = self._buf("noise.actor", H, N, d.A).exponential_()
        else:
            actor = self._buf("noise.actor", H, N, d.A).normal_()
        prior = self._buf("noise.prior", H, N, d.Z).exponential_()
        out = dict(post=post, actor=actor, prior=prior)
        if image_pred:
            out["image_pred"] = self._buf("noise.image_pred", N, d.Z).exponential_()
        if dream_log:
            la = self._buf("noise.dl_actor", T - 1, B, d.A)
            out["dream_log_actor"] = la.exponential_() if self.conf.actor_dist == "onehot" else la.normal_()
            out["dream_log_prior"] = self._buf("noise.dl_prior", T - 1, B, d.Z).exponential_()
        return out

    # ------------------------------------------------------------------ training step
    def training_step(self, obs, in_state, iwae_samples=None, imag_horizon=None, do_open_loop=False,
                      do_image_pred=False, do_dream_tensors=False, noise=None):
        """dreamer.py:113-186.  `noise` (optional, tests): dict(post=(T,BI,Z) Exp(1), actor=(H,N,A),
        prior=(H,N,Z) Exp(1)) replacing the internally drawn sampling noise."""
        assert "action" in obs, "`action` required in observation"
        assert "reward" in obs, "`reward` required in observation"
        assert "reset" in obs, "`reset` required in observation"
        assert "terminal" in obs, "`terminal` required in observation"
        I = int(iwae_samples or self.iwae_samples)
        H = int(imag_horizon or self.imag_horizon)
        T, B = obs["action"].shape[:2]
        self._ensure_arena()
        want_grad = torch.is_grad_enabled()
        with torch.no_grad():
            if want_grad:
                self._note_new_grads()
                self._sync_target_critic()
            flags = (bool(do_open_loop), bool(do_image_pred), bool(do_dream_tensors and self.wm.decoder.image is not None))
            if flags[0] and want_grad:
                raise NotImplementedError("do_open_loop is the evaluation branch (train.py:353-359 runs it under no_grad)")
            graphed = self.use_cuda_graph and noise is None and want_grad and self._arena.is_cuda and not any(flags)
            if graphed:
                wm_out, ac_out = self._graphed_core(obs, in_state, T, B, I, H)
            else:                                   # logging / evaluation steps (~10 % of steps) are launched eagerly
                wm_out, ac_out = self._core(obs, in_state, T, B, I, H, noise, want_grad, flags=flags)
        loss_model, loss_probe = wm_out["loss_model"], self.probe_model.dummy.detach() ** 2
        loss_actor, loss_critic = ac_out["loss_actor"], ac_out["loss_critic"]
        if want_grad:
            gp = self._group_params
            loss_model = _AttachGrads.apply(loss_model, self, "wm", *gp["wm"])
            loss_probe = _AttachGrads.apply(loss_probe, self, "probe", *gp["probe"])
            loss_actor = _AttachGrads.apply(loss_actor, self, "actor", *gp["actor"])
            loss_critic = _AttachGrads.apply(loss_critic, self, "critic", *gp["critic"])
        metrics = dict(wm_out["metrics"]); metrics.update(ac_out["metrics"])
        # the step's scalars live in reused workspace / CUDA-graph buffers: hand the caller its own copy (one small gather
        # kernel), so metrics kept across steps (train.py:204-214 accumulates them for logging) stay what they were
        keys = list(metrics)
        snap = torch.stack([metrics[k].detach().reshape(()).to(torch.float32) for k in keys])
        metrics = {k: snap[i] for i, k in enumerate(keys)}
        tensors = dict(wm_out["tensors"])
        tensors.update(policy_value=ac_out["value"][0].reshape(T, B, I).mean(-1))
        if not graphed:
            # logging / evaluation steps (the ones whose tensors train.py actually reads) return private copies; on the
            # CUDA-graph steady-state path `tensors` stay views of the step workspace, valid until the next call
            tensors = {k: v.clone() for k, v in tensors.items()}
        return (loss_model, loss_probe, loss_actor, loss_critic), wm_out["out_state"], metrics, tensors, \
            ac_out.get("dream_tensors", {})

    def _core(self, obs, in_state, T, B, I, H, noise, want_grad, force_weights=False, flags=(False, False, False)):
        """The kernel schedule of one step: weights prep, WM forward (+backward), dream, actor-critic (+backward)."""
        if force_weights:
            self._weights_dirty = True
        self._prepare_weights()
        N = T * B * I
        open_loop, image_pred, dream_log = flags
        if noise is None:
            noise = self._draw_noise(T, B * I, N, H, image_pred, dream_log, B)
        if want_grad:
            self.ops.fill(self._garena, 0.0)
        tm = self._phase_timer
        if tm is not None:
            tm.mark("prepare+noise")
        par = self._ov(1)
        ac_box = {}

        def run_ac():
            feats = self._buf("feats", H + 1, N, self.d.F)
            self._dream(feats, N, H, noise["actor"], noise["prior"], "")
            if tm is not None:
                tm.mark("dream")
            ac_box["out"] = self._actor_critic(feats, N, H, want_grad, "")
            if tm is not None:
                tm.mark("actor_critic")

        def after_features():                       # the dream needs only the (detached) posterior features
            if par:
                self._scratch_ns = "ac."
                try:
                    with self._fork(1):
                        run_ac()
                finally:
                    self._scratch_ns = ""

        wm_out = self._wm_forward(obs, in_state, T, B, I, H, noise["post"], open_loop,
                                  noise["image_pred"] if image_pred else None, after_features=after_features)
        if tm is not None:
            tm.mark("wm_forward")
        if want_grad:
            self._wm_backward(obs, T, B, I, H)
        if tm is not None:
            tm.mark("wm_backward")
        if par:
            self._join(1)
            cur = torch.cuda.current_stream(self._arena.device)
            for v in ac_box["out"]["metrics"].values():       # allocated on the side stream, consumed on this one
                v.record_stream(cur)
        else:
            run_ac()
        ac_out = ac_box["out"]
        if dream_log:
            ac_out["dream_tensors"] = self._dream_for_log(obs, T, B, I, noise["dream_log_actor"], noise["dream_log_prior"])
        return wm_out, ac_out

    _phase_timer = None       # bench.py installs a PhaseTimer (CUDA events between the phases of one eager step)

    def _sync_target_critic(self):
        """a2c.py:76-79: copy critic -> critic_target every target_interval calls (host-side counter)."""
        ac = self.ac
        if ac.train_steps % self.conf.target_interval == 0:
            self._group_slice("target", self._arena).copy_(self._group_slice("critic", self._arena))
            self._weights_dirty = True
        ac.train_steps += 1

    # set PD_B200_GRAPHS=0 to launch every kernel from Python instead of replaying a captured CUDA graph
    use_cuda_graph = os.environ.get("PD_B200_GRAPHS", "1") != "0"

    # Independent parts of the step are issued on side streams (parallel branches of the captured graph), so the
    # latency-bound M=B recurrent chains share the GPU with throughput-bound work instead of idling it.  Bit mask:
    #   1  dream + actor-critic (needs only the detached features) alongside decoder / losses / world-model backward
    #   2  the h·W_hh GEMM of the next step (and its transpose in BPTT) off the per-timestep critical path
    #   4  image-decoder weight gradients alongside the input-gradient chain and BPTT
    # (Measured, profiles/README.md: 1 is worth 4.5 ms of 36; 2 and 4 are within noise; running the encoder in time
    #  chunks beside the unroll / BPTT gained nothing - the split-K chain GEMMs already occupy every SM.)
    overlap = int(os.environ.get("PD_B200_OVERLAP", "3"))
    _scratch_ns = ""          # name space of the shared MLP scratch buffers (one per concurrent branch)

    # The posterior unroll runs as ONE cooperative kernel (csrc/pd_rssm_fwd3.cu) when the shape fits its limits
    # (B*I <= 256 rows, ...); PD_B200_PERSISTENT_RSSM=0 selects the chain of 9 launches per timestep instead.
    persistent_rssm = os.environ.get("PD_B200_PERSISTENT_RSSM", "1") != "0"

    # BPTT through the posterior unroll as ONE cooperative kernel (csrc/pd_rssm_bptt.cu): opt-in with PD_B200_PERSISTENT_BPTT=1.
    # Default is the chain of ~12 launches per timestep: standalone the two take the same 4 ms, but the chain's latency-bound
    # launches share the SMs with the concurrent imagination branch while a cooperative kernel owns all of them for its whole
    # duration (measured on B200, r02: 27.7 ms/step with the chain, 30.6 ms with the kernel).
    persistent_bptt = os.environ.get("PD_B200_PERSISTENT_BPTT", "0") != "0"

    def _persistent_bptt_ok(self, BI):
        d = self.d
        on_gpu = self._arena.is_cuda
        if not (self.persistent_bptt and self._dp_allows() and (on_gpu or self.ops.is_reference) and
                getattr(self, "_k1b_w", None)):
            return False
        P = torch.cuda.get_device_properties(self._arena.device).multi_processor_count if on_gpu else 148
        Z = d.G * d.C
        ks2 = 4 if Z % 256 == 0 and P >= 4 else 1
        ks6 = 4 if (3 * d.D) % 256 == 0 and P >= 4 else 1
        R = max(1, min(4, P // d.G))
        cd = lambda a_, b_: -(-a_ // b_)
        return (BI <= min(64, P) and d.Hd <= 1024 and d.Hd % 8 == 0 and d.D % 8 == 0 and Z % 8 == 0 and d.C <= 32 and d.G <= P and
                cd(BI, R) <= 16 and cd(d.D, P) <= 16 and cd(d.Hd, P // ks2) <= 32 and cd(d.D, P // ks6) <= 64 and
                cd(d.Hd, P // ks6) <= 32)

    def _persistent_rssm_ok(self, BI):
        d = self.d
        on_gpu = self._arena.is_cuda
        if not (self.persistent_rssm and self.fp16_forward and self._dp_allows() and (on_gpu or self.ops.is_reference)):
            return False
        P = torch.cuda.get_device_properties(self._arena.device).multi_processor_count if on_gpu else 148
        ks = 4 if d.D % 256 == 0 and P >= 4 else 1
        R = max(1, min(4, P // d.G))
        cd = lambda a_, b_: -(-a_ // b_)
        # batch rows (B x iwae_samples) beyond one 64-row MMA operand are taken in blocks by the kernel, up to 256
        return (BI <= 256 and d.Hd <= 1024 and d.Hd % 8 == 0 and d.D % 8 == 0 and d.C <= 32 and d.G <= P and
                cd(d.D, P) <= 16 and cd(d.D, P // ks) <= 64 and cd(d.Hd, P // ks) <= 32 and
                getattr(self, "_k1_wzT", None) is not None)

    def _ov(self, bit):
        # (the eager phase timer of bench.py needs one stream)
        return bool(self.overlap & bit) and self._arena.is_cuda and self._phase_timer is None and self._dp_allows()

    def _dp_allows(self):
        # Data-parallel runs use the SAME schedule as one GPU (side-stream branches, persistent RSSM kernels, graph replay):
        # the r02 two-GPU triage matrix (profiles/r02_dp_triage.md) completed in every combination.  PD_B200_DP_FEATURES=0
        # restores round 1's conservative single-stream / per-timestep-chain schedule under data parallelism.
        return self._dp is None or os.environ.get("PD_B200_DP_FEATURES", "1") != "0"

    def _side(self, k):
        key = (k, torch.cuda.current_stream(self._arena.device).cuda_stream)     # one side stream per (purpose, parent)
        st = self.__dict__.setdefault("_side_streams", {})
        if key not in st:
            st[key] = torch.cuda.Stream(device=self._arena.device)
        return st[key]

    def _fork(self, k):
        """Context manager: the block is issued on side stream k, ordered after everything issued so far."""
        s = self._side(k)
        s.wait_stream(torch.cuda.current_stream(self._arena.device))
        return torch.cuda.stream(s)

    def _join(self, k):
        torch.cuda.current_stream(self._arena.device).wait_stream(self._side(k))

    def _graphed_core(self, obs, in_state, T, B, I, H):
        """CUDA-graph replay of `_core` (the ~1 400 kernel launches of a step are issued by one cudaGraphLaunch).
        Calls 1-2 for a shape run eagerly (allocates the workspace, loads the kernels); call 3 captures."""
        key = (T, B, I, H) + tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(obs.items()))
        st = self._graphs.setdefault(key, dict(calls=0, graph=None))
        st["calls"] += 1
        if st["graph"] is None and (st["calls"] <= 2 or st.get("failed")):
            return self._core(obs, in_state, T, B, I, H, None, True)
        if st["graph"] is None:
            st["obs"] = {k: torch.empty_like(v) for k, v in obs.items()}
            st["state"] = tuple(torch.empty_like(s_) for s_ in in_state)
            for k, v in obs.items():
                st["obs"][k].copy_(v)
            for d_, s_ in zip(st["state"], in_state):
                d_.copy_(s_)
            torch.cuda.synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                k0 = self.ops.launch_count()
                with torch.cuda.graph(g):
                    st["out"] = self._core(st["obs"], st["state"], T, B, I, H, None, True, force_weights=True)
                st["kernels"] = self.ops.launch_count() - k0      # kernel nodes of this library in the graph
                st["graph"] = g
            except Exception as e:                           # keep running eagerly (same kernels), say so once
                st["failed"] = True
                import warnings
                warnings.warn(f"pydreamer_b200: CUDA graph capture failed ({e}); continuing with eager launches")
                torch.cuda.synchronize()
                return self._core(obs, in_state, T, B, I, H, None, True)
        else:
            for k, v in obs.items():
                st["obs"][k].copy_(v, non_blocking=True)
            for d_, s_ in zip(st["state"], in_state):
                d_.copy_(s_, non_blocking=True)
        st["graph"].replay()
        return st["out"]

    # ------------------------------------------------------------------ world model forward
    def _wm_features(self, obs, in_state, T, B, I, noise_post, tag, feat, open_loop=False):
        """Encoder + posterior unroll (forward only).  `feat` (T, B*I, F) receives cat(h, z); every intermediate the
        backward needs is kept in workspace buffers named `tag + ...`."""
        ops, d, conf = self.ops, self.d, self.conf
        NB, BI = T * B, B * I
        N = NB * I
        cd, IC = d.cd, d.IC
        b = lambda name, *shape, **kw: self._buf(tag + name, *shape, **kw)
        enc = self.wm.encoder.encoder_image.model
        cell = self.wm.core.cell
        gru = cell.gru.layers[0]
        # ---- encoder (encoders.py:72-96): im2col -> tcgen05 GEMM (+bias+ELU) x4, NHWC activations
        img = obs["image"].reshape(NB, IC, 64, 64)
        geo = ((64, 31, IC, cd), (31, 14, cd, 2 * cd), (14, 6, 2 * cd, 4 * cd), (6, 2, 4 * cd, 8 * cd))
        embed = b("enc.embed", NB, d.E)
        ea = b("rssm.ea", NB, d.Hd)

        def encode(r0, r1):                         # images [r0, r1) -> embed rows -> hoisted post_mlp_e product
            x4 = img[r0:r1].permute(0, 2, 3, 1)
            for li, (hin_, hout, ci, co) in enumerate(geo):
                hw = hout * hout
                act = b(f"enc.a{li}", NB * hw, co)[r0 * hw:r1 * hw]
                if li > 0 and self.implicit_conv:
                    ops.conv_gemm(1, x4, 4, self._encw[li], act, bias=self._raw(enc[2 * li].bias), act=ACT_ELU,
                                  round_out=True)
                else:
                    col = b(f"enc.col{li}", NB * hw, 16 * ci)[r0 * hw:r1 * hw]
                    ops.im2col(x4, 4, 1 if li == 0 else 0, col, round_out=True)
                    ops.gemm(col, self._encw[li], act, bias=self._raw(enc[2 * li].bias), act=ACT_ELU, round_out=True)
                x4 = act.view(r1 - r0, hout, hout, co)
            ops.permute4(x4.view(r1 - r0, 4, 8 * cd, 1), embed[r0:r1].view(r1 - r0, 8 * cd, 4, 1), (0, 2, 1, 3),
                         round_out=True)                                       # (h,w,c) -> reference (c,h,w) flatten
            ops.gemm(embed[r0:r1], self._w(cell.post_mlp_e.weight), ea[r0:r1])  # hoisted over T

        encode(0, NB)

        # ---- RSSM posterior unroll (rssm.py:21-78, 125-153)
        mask = b("rssm.mask", T, BI)
        ops.reset_mask(obs["reset"], I, mask)
        action = obs["action"].reshape(NB, d.A)
        aa = b("rssm.aa", NB, d.Hd); ops.gemm(action, self._w(cell.a_mlp.weight), aa)
        hin, zin = b("rssm.hin", T, BI, d.D), b("rssm.zin", T, BI, d.Z)
        ops.mask_rows(in_state[0], mask[0], hin[0]); ops.mask_rows(in_state[1], mask[0], zin[0])
        x1, za = b("rssm.x1", T, BI, d.Hd), b("rssm.za", T, BI, d.Hd)
        m1, r1 = b("rssm.m1", T, BI), b("rssm.r1", T, BI)
        gates = b("rssm.gates", T, BI, 4 * d.D)
        y2, pin = b("rssm.y2", T, BI, d.Hd), b("rssm.pin", T, BI, d.Hd)
        m2, r2 = b("rssm.m2", T, BI), b("rssm.r2", T, BI)
        post = b("rssm.post", T, BI, d.Z)
        idx = b("rssm.idx", T, BI, d.G, dtype=torch.int32)
        W = self._w
        if self._persistent_rssm_ok(BI):
            try:
                # one cooperative kernel for all T steps (csrc/pd_rssm_fwd3.cu); step 0's pre-norm input is formed
                # here because the incoming z need not be one-hot
                Wh, h16 = self._wh, torch.float16
                ops.gemm(zin[0], W(cell.z_mlp.weight), x1[0], bias=self._raw(cell.z_mlp.bias), res=aa[:B], r_div=I)
                ph, pn, pm = ((cell.prior_mlp_h, cell.prior_norm, cell.prior_mlp) if open_loop else
                              (cell.post_mlp_h, cell.post_norm, cell.post_mlp))
                ops.rssm_unroll_fwd(
                    dict(T=T, BI=BI, I=I, D=d.D, Hd=d.Hd, G=d.G, C=d.C), 1e-3,
                    w_z16=Wh(cell.z_mlp.weight), w_ih16=Wh(gru.weight_ih), w_hh16=Wh(gru.weight_hh), w_ph16=Wh(ph.weight),
                    w_pm16=Wh(pm.weight), b_z=self._raw(cell.z_mlp.bias), ln1_g=self._raw(cell.in_norm.weight),
                    ln1_b=self._raw(cell.in_norm.bias), b_ih=self._raw(gru.bias_ih), b_hh=self._raw(gru.bias_hh),
                    b_ph=self._raw(ph.bias), ln2_g=self._raw(pn.weight), ln2_b=self._raw(pn.bias), b_pm=self._raw(pm.bias),
                    aa=aa, ea=None if open_loop else ea, mask=mask, noise=noise_post, x1=x1, za=za, m1=m1, r1=r1,
                    gates=gates, feat=feat, hin=hin, zin=zin, y2=y2, pin=pin, m2=m2, r2=r2, post=post, idx=idx,
                    ws_wzT16=self._k1_wzT, ws_za16=b("k1.za16", BI, d.Hd, dtype=h16),
                    ws_h16=b("k1.h16", BI, d.D, dtype=h16), ws_pin16=b("k1.pin16", BI, d.Hd, dtype=h16),
                    ws_barrier=b("k1.bar", 16, dtype=torch.int32), ws_ghpart=b("k1.ghpart", 4, BI, 3 * d.D),
                    ws_y2part=b("k1.y2part", 4, BI, d.Hd))
                out_state = (feat[T - 1, :, :d.D].clone(), feat[T - 1, :, d.D:].clone())
                return img, post, idx, out_state
            except RuntimeError as e:        # e.g. cooperative launch refused (SMs reserved by MPS / green contexts)
                import warnings
                warnings.warn(f"pydreamer_b200: persistent RSSM kernel unavailable ({e}); using the per-timestep chain")
                self.persistent_rssm = False

        gi, gh = b("rssm.gi", T, BI, 3 * d.D), b("rssm.gh", T, BI, 3 * d.D)
        skinny = BI <= 128                      # the per-timestep GEMMs split K and reduce into C: clear all T slices at once
        if skinny:
            for buf_ in (x1, gi, gh, y2, post):
                ops.fill(buf_, 0.0)
        par = self._ov(2)
        gh_gemm = lambda t: ops.gemm(hin[t], W(gru.weight_hh), gh[t], bias=self._raw(gru.bias_hh), c_zeroed=skinny)
        if par:
            with self._fork(2):
                gh_gemm(0)
        for t in range(T):
            last = t == T - 1
            ops.gemm(zin[t], W(cell.z_mlp.weight), x1[t], bias=self._raw(cell.z_mlp.bias), res=aa[t * B:(t + 1) * B],
                     r_div=I, c_zeroed=skinny)
            ops.ln_elu_fwd(x1[t], self._raw(cell.in_norm.weight), self._raw(cell.in_norm.bias), 1e-3, za[t], m1[t], r1[t])
            ops.gemm(za[t], W(gru.weight_ih), gi[t], bias=self._raw(gru.bias_ih), c_zeroed=skinny)
            if par:
                self._join(2)
            else:
                gh_gemm(t)
            ops.gru_fwd(gi[t], gh[t], hin[t], feat[t, :, :d.D], None if last else hin[t + 1],
                        None if last else mask[t + 1], gates[t])
            if par and not last:                    # h_{t+1} is known: its W_hh product overlaps the posterior MLP
                with self._fork(2):
                    gh_gemm(t + 1)
            if not open_loop:
                ops.gemm(feat[t, :, :d.D], W(cell.post_mlp_h.weight), y2[t], bias=self._raw(cell.post_mlp_h.bias),
                         res=ea[t * B:(t + 1) * B], r_div=I, c_zeroed=skinny)
                ops.ln_elu_fwd(y2[t], self._raw(cell.post_norm.weight), self._raw(cell.post_norm.bias), 1e-3, pin[t],
                               m2[t], r2[t])
                ops.gemm(pin[t], W(cell.post_mlp.weight), post[t], bias=self._raw(cell.post_mlp.bias), c_zeroed=skinny)
            else:                                   # open loop (rssm.py:52-53): the "posterior" is the prior, no embed
                ops.gemm(feat[t, :, :d.D], W(cell.prior_mlp_h.weight), y2[t], bias=self._raw(cell.prior_mlp_h.bias),
                         c_zeroed=skinny)
                ops.ln_elu_fwd(y2[t], self._raw(cell.prior_norm.weight), self._raw(cell.prior_norm.bias), 1e-3, pin[t],
                               m2[t], r2[t])
                ops.gemm(pin[t], W(cell.prior_mlp.weight), post[t], bias=self._raw(cell.prior_mlp.bias), c_zeroed=skinny)
            ops.cat_sample(post[t], noise_post[t], d.G, d.C, feat[t, :, d.D:], None if last else zin[t + 1],
                           None if last else mask[t + 1], idx[t])
        out_state = (feat[T - 1, :, :d.D].clone(), feat[T - 1, :, d.D:].clone())
        return img, post, idx, out_state

    def _wm_forward(self, obs, in_state, T, B, I, H, noise_post, open_loop=False, noise_image_pred=None,
                    after_features=None):
        ops, d, conf = self.ops, self.d, self.conf
        NB, BI = T * B, B * I
        N = NB * I
        b, W = self._buf, self._w
        cell = self.wm.core.cell
        feats = b("feats", H + 1, N, d.F)            # feats[0] = world-model features, feats[1:] = dream
        img, post, idx, out_state = self._wm_features(obs, in_state, T, B, I, noise_post, "", feats[0].view(T, BI, d.F),
                                                      open_loop)
        if after_features is not None:
            after_features()
        featN = feats[0]                                   # (N, F)
        hN = featN[:, :d.D]
        # batched prior (rssm.py:186-193)
        yp, ppin = b("rssm.yp", N, d.Hd), b("rssm.ppin", N, d.Hd)
        m3, r3 = b("rssm.m3", N), b("rssm.r3", N)
        prior = b("rssm.prior", N, d.Z)
        ops.gemm(hN, W(cell.prior_mlp_h.weight), yp, bias=self._raw(cell.prior_mlp_h.bias))
        ops.ln_elu_fwd(yp, self._raw(cell.prior_norm.weight), self._raw(cell.prior_norm.bias), 1e-3, ppin, m3, r3)
        ops.gemm(ppin, W(cell.prior_mlp.weight), prior, bias=self._raw(cell.prior_mlp.bias))

        # ---- image decoder + reward / terminal heads on the posterior features
        dd = self._decode_all(featN, img, obs, N, NB, I, "")
        image_dec, l_img, l_rew, l_term, rec_r, rec_t = dd["image"], dd["l_img"], dd["l_rew"], dd["l_term"], dd["rec_r"], dd["rec_t"]

        # ---- KL + loss assembly (dreamer.py:328-379)
        l_kl, kl_exact = b("loss.kl", N), b("loss.klx", N)
        ent_post, ent_prior = b("loss.entq", N), b("loss.entp", N)
        dpost_u, dprior = b("kl.dpost", N, d.Z), b("kl.dprior", N, d.Z)
        kb = conf.kl_balance
        ops.kl(post.view(N, d.Z), prior, idx.view(N, d.G), 0 if I == 1 else 1, -1.0 if kb == 0.5 else kb, d.G, d.C,
               l_kl, kl_exact, ent_post, ent_prior, dpost_u, dprior)
        w, tb = b("loss.w", N), b("loss.tb", NB, 8)
        ops.wm_loss(NB, I, conf.kl_weight, conf.image_weight, conf.reward_weight, conf.terminal_weight, l_img, l_rew,
                    l_term, l_kl, kl_exact, ent_prior, ent_post, w, tb)
        means = b("loss.means", 8)
        ops.colmean(tb, means)
        tbv = tb.view(T, B, 8)
        metrics = dict(loss_image=means[1], loss_reward=means[2], loss_terminal=means[3], loss_model=means[0],
                       loss_kl=means[4], entropy_prior=means[5], entropy_post=means[6])
        sel = (lambda x: x.view(T, B, I, *x.shape[1:])[:, :, 0]) if I == 1 else \
              (lambda x: x.view(T, B, I, *x.shape[1:]).mean(2))
        tensors = dict(loss_image=tbv[..., 1], image_rec=sel(image_dec), loss_reward=tbv[..., 2],
                       reward_rec=sel(rec_r), loss_terminal=tbv[..., 3], terminal_rec=sel(rec_t),
                       loss_kl=tbv[..., 4], entropy_prior=tbv[..., 5], entropy_post=tbv[..., 6])
        if noise_image_pred is not None:
            self._image_pred(obs, featN, prior, noise_image_pred, T, B, I, metrics, tensors)
        return dict(loss_model=means[0], out_state=out_state, metrics=metrics, tensors=tensors)

    def _image_pred(self, obs, featN, prior, noise, T, B, I, metrics, tensors):
        """dreamer.py:383-394: decode from a PRIOR sample (what the model predicts before seeing the observation);
        reports the reconstruction losses as logprob_* and the decoded tensors as *_pred.  Logging branch."""
        ops, d = self.ops, self.d
        NB = T * B
        N = NB * I
        b = self._buf
        featP = b("p.feat", N, d.F)
        featP[:, :d.D].copy_(featN[:, :d.D])
        ops.cat_sample(prior, noise, d.G, d.C, featP[:, d.D:])
        img = obs["image"].reshape(NB, d.IC, 64, 64)
        dd = self._decode_all(featP, img, obs, N, NB, I, "p.")
        zeros = b("p.zeros", N, zero=True)
        w, tb = b("p.loss.w", N), b("p.loss.tb", NB, 8)
        ops.wm_loss(NB, I, 0.0, 1.0, 1.0, 1.0, dd["l_img"], dd["l_rew"], dd["l_term"], zeros, zeros, zeros, zeros, w, tb)
        tbv = tb.view(T, B, 8)
        sel = (lambda x: x.view(T, B, I, *x.shape[1:])[:, :, 0]) if I == 1 else \
              (lambda x: x.view(T, B, I, *x.shape[1:]).mean(2))
        lp_img, lp_rew, lp_term = tbv[..., 1], tbv[..., 2], tbv[..., 3]
        nanmean = lambda x: torch.nansum(x) / (~torch.isnan(x)).sum()                # functions.py:150-151
        extra_t = {}
        for sig in (-1, 1):                                                          # decoders.py:96-101
            m = torch.sign(obs["reward"]) == sig
            extra_t[f"logprob_reward{sig}"] = lp_rew * m / m
        m = obs["terminal"] > 0                                                      # decoders.py:103-106
        extra_t["logprob_terminal1"] = lp_term * m / m
        metrics.update(logprob_image=lp_img.mean(), logprob_reward=lp_rew.mean(), logprob_terminal=lp_term.mean(),
                       **{k: nanmean(v) for k, v in extra_t.items()})
        tensors.update(logprob_image=lp_img, logprob_reward=lp_rew, logprob_terminal=lp_term, **extra_t,
                       image_pred=sel(dd["image"]), reward_pred=sel(dd["rec_r"]), terminal_pred=sel(dd["rec_t"]))

    def _cols_dtype(self, ncols):
        """Column matrices of the transposed convolutions are written once and read once: fp16 halves that traffic.  The GEMM's
        fp16 TMA store needs 16-byte rows (ncols % 8 == 0); the last layer (k*k*3 columns) stays fp32."""
        return torch.float16 if (self.fp16_forward and self.fp16_cols and ncols % 8 == 0) else torch.float32

    def _decode_all(self, featN, img, obs, N, NB, I, tag):
        """MultiDecoder.training_step forward (decoders.py:50-108) on features (N,F): image decoder (Linear, then each
        deconv = GEMM + col2im gather with bias+ELU; the last one fused with the image loss) and the reward / terminal
        MLP heads with their losses.  Buffers are named `tag + ...` (tag "" = the training pass the backward reads)."""
        ops, d = self.ops, self.d
        cd, IC = d.cd, d.IC
        b = lambda name, *shape, **kw: self._buf(tag + name, *shape, **kw)
        W = self._w
        dec = self.wm.decoder.image.model
        x0 = b("dec.x0", N, 32 * cd)
        ops.gemm(featN, W(dec[0].weight), x0, bias=self._raw(dec[0].bias), round_out=True)
        dgeo = ((1, 5, 5, 32 * cd, 4 * cd), (5, 13, 5, 4 * cd, 2 * cd), (13, 30, 6, 2 * cd, cd), (30, 64, 6, cd, IC))
        xin = x0
        for li, (hi, ho, k, ci, co) in enumerate(dgeo):
            cols = b(f"dec.cols{li}", N * hi * hi, k * k * co, dtype=self._cols_dtype(k * k * co))
            ops.gemm(xin, self._decw[li], cols)
            bias = self._raw(dec[2 + 2 * li].bias)
            if li < 3:
                a = b(f"dec.d{li}", N, ho, ho, co)
                ops.col2im(cols, hi, hi, k, bias, ACT_ELU, a, round_out=True)
                xin = a.view(N * ho * ho, co)
            else:
                image_dec, diff = b("dec.image", N, IC, 64, 64), b("dec.diff", N, IC, 64, 64)
                l_img, csum = b("loss.img", N), b("dec.csum", N, IC)
                ops.col2im_imgloss(cols, N, hi, hi, IC, k, bias, img, I, image_dec, diff, l_img, csum)
        # reward / terminal heads (decoders.py:257-319)
        rp, tp = self._mlp_params(self.wm.decoder.reward.model), self._mlp_params(self.wm.decoder.terminal.model)
        yr, yt = b("head.yr", N, 1), b("head.yt", N, 1)
        self._mlp_fwd(rp, featN, yr, tag + "rew", save=True)
        self._mlp_fwd(tp, featN, yt, tag + "term", save=True)
        l_rew, dyr, rec_r = b("loss.rew", N), b("head.dyr", N, 1), b("head.rec_r", N)
        l_term, dyt, rec_t = b("loss.term", N), b("head.dyt", N, 1), b("head.rec_t", N)
        ops.scalar_head_loss(0, yr, obs["reward"].reshape(NB), I, l_rew, dyr, rec_r)
        ops.scalar_head_loss(1, yt, obs["terminal"].reshape(NB), I, l_term, dyt, rec_t)
        return dict(image=image_dec, l_img=l_img, l_rew=l_rew, l_term=l_term, rec_r=rec_r, rec_t=rec_t)

    # ------------------------------------------------------------------ world model backward
    def _wm_backward(self, obs, T, B, I, H):
        ops, d, conf = self.ops, self.d, self.conf
        NB, BI = T * B, B * I
        N = NB * I
        cd, IC = d.cd, d.IC
        b, W, G = self._buf, self._w, self._g
        cell = self.wm.core.cell
        gru = cell.gru.layers[0]
        enc = self.wm.encoder.encoder_image.model
        dec = self.wm.decoder.image.model
        featN = b("feats", H + 1, N, d.F)[0]
        w = b("loss.w", N)
        dfeat = b("bwd.dfeat", N, d.F)

        # ---- image decoder backward: seeds = w[n] * image_weight * (dec - target)
        diff = b("dec.diff", N, IC, 64, 64)
        csum = b("dec.csum", N, IC)
        ops.rowscale(diff.view(N, IC * 4096), w, 1, conf.image_weight)
        ops.rowscale(csum, w, 1, conf.image_weight)
        ops.colsum(csum, G(dec[8].bias))
        dgeo = ((1, 5, 5, 32 * cd, 4 * cd), (5, 13, 5, 4 * cd, 2 * cd), (13, 30, 6, 2 * cd, cd), (30, 64, 6, cd, IC))
        impl = [self.implicit_conv and li in (1, 2) for li in range(4)]   # deconv 2,3: 32-channel-aligned NHWC gradients
        copad = [(dgeo[li][4] + 31) // 32 * 32 for li in range(4)]
        gdec = [b(f"bwd.gdecwp{li}", dgeo[li][2] ** 2 * copad[li], dgeo[li][3]) if impl[li]
                else b(f"bwd.gdecw{li}", *self._decw[li].shape) for li in range(4)]
        for g_ in gdec:
            ops.fill(g_, 0.0)
        par_w = self._ov(4)                 # weight gradients leave the dfeat -> BPTT critical path (joined at the end)
        side = (lambda: self._fork(4)) if par_w else contextlib.nullcontext
        dout4 = diff.permute(0, 2, 3, 1)                       # [n,y,x,c] view of the NCHW diff
        for li in (3, 2, 1, 0):
            hi, ho, k, ci, co = dgeo[li]
            xin = b("dec.x0", N, 32 * cd) if li == 0 else b(f"dec.d{li - 1}", N, hi, hi, ci).view(N * hi * hi, ci)
            dxin = b(f"bwd.dd{li}", N * hi * hi, ci)
            # (li > 0: the ELU backward and the bias gradient of the deconv below ride in the input-gradient GEMM's epilogue)
            below_bias = G(dec[2 * li].bias) if li > 0 else None
            if impl[li]:
                with side():
                    ops.conv_gemm(2, dout4, k, xin, gdec[li])                          # weight gradient, rows (tap, co padded)
                ops.conv_gemm_actbwd(dout4, k, self._decw[li], dxin, xin, below_bias, o_mn=True)   # input gradient
            else:
                if li == 0:
                    dcols = dout4.reshape(N, k * k * co)       # 5x5 input of a 5x5 kernel: im2col is the identity
                else:
                    dcols = b(f"bwd.dcols{li}", N * hi * hi, k * k * co)
                    ops.im2col(dout4, k, 0, dcols, round_out=True)
                with side():
                    ops.gemm(dcols, xin, gdec[li], a_mn=True, b_mn=True, accumulate=True)
                if li > 0:
                    ops.gemm_actbwd(dcols, self._decw[li], dxin, xin, below_bias, b_mn=True)
                else:
                    ops.gemm(dcols, self._decw[li], dxin, b_mn=True, round_out=True)
            if li > 0:
                dout4 = dxin.view(N, hi, hi, ci)
            else:
                dx0 = dxin
        with side():
            for li, idx_ in enumerate((2, 4, 6, 8)):          # back to ConvTranspose2d layout (Cin,Cout,kh,kw)
                wt = dec[idx_].weight
                ci, co, kh, kw = wt.shape
                src = gdec[li].view(kh, kw, copad[li], ci)[:, :, :co] if impl[li] else gdec[li].view(kh, kw, co, ci)
                ops.permute4(src, G(wt), (3, 2, 0, 1))
            ops.gemm(dx0, featN, G(dec[0].weight), a_mn=True, b_mn=True, accumulate=True)
            ops.colsum(dx0, G(dec[0].bias))
        ops.gemm(dx0, W(dec[0].weight), dfeat, b_mn=True)                      # first writer of dfeat

        # ---- reward / terminal heads
        rp, tp = self._mlp_params(self.wm.decoder.reward.model), self._mlp_params(self.wm.decoder.terminal.model)
        dyr, dyt = b("head.dyr", N, 1), b("head.dyt", N, 1)
        ops.rowscale(dyr, w, 1, conf.reward_weight)
        ops.rowscale(dyt, w, 1, conf.terminal_weight)
        self._mlp_bwd(rp, featN, dyr, "rew", din=dfeat, din_accum=True)
        self._mlp_bwd(tp, featN, dyt, "term", din=dfeat, din_accum=True)

        # ---- prior branch (batch_prior): dprior = kl_weight * w[n] * dKL/dprior
        dprior = b("kl.dprior", N, d.Z)
        ops.rowscale(dprior, w, 1, conf.kl_weight)
        ppin, yp = b("rssm.ppin", N, d.Hd), b("rssm.yp", N, d.Hd)
        dpp, dyp = b("bwd.dpp", N, d.Hd), b("bwd.dyp", N, d.Hd)
        ops.gemm(dprior, ppin, G(cell.prior_mlp.weight), a_mn=True, b_mn=True, accumulate=True)
        ops.colsum(dprior, G(cell.prior_mlp.bias))
        ops.gemm(dprior, W(cell.prior_mlp.weight), dpp, b_mn=True)
        ops.ln_elu_bwd(dpp, yp, ppin, self._raw(cell.prior_norm.weight), b("rssm.m3", N), b("rssm.r3", N), dyp,
                       G(cell.prior_norm.weight), G(cell.prior_norm.bias), G(cell.prior_mlp_h.bias))
        ops.gemm(dyp, featN[:, :d.D], G(cell.prior_mlp_h.weight), a_mn=True, b_mn=True, accumulate=True)
        ops.gemm(dyp, W(cell.prior_mlp_h.weight), dfeat[:, :d.D], b_mn=True, res=dfeat[:, :d.D])

        # ---- BPTT through the posterior unroll
        dfeat3 = dfeat.view(T, BI, d.F)
        mask = b("rssm.mask", T, BI)
        post, pin, y2 = b("rssm.post", T, BI, d.Z), b("rssm.pin", T, BI, d.Hd), b("rssm.y2", T, BI, d.Hd)
        m2, r2 = b("rssm.m2", T, BI), b("rssm.r2", T, BI)
        x1, za = b("rssm.x1", T, BI, d.Hd), b("rssm.za", T, BI, d.Hd)
        m1, r1 = b("rssm.m1", T, BI), b("rssm.r1", T, BI)
        gates, hin, zin = b("rssm.gates", T, BI, 4 * d.D), b("rssm.hin", T, BI, d.D), b("rssm.zin", T, BI, d.Z)
        dpost_u = b("kl.dpost", N, d.Z).view(T, BI, d.Z)
        w3 = w.view(T, BI)
        dpost = b("bwd.dpost", T, BI, d.Z)
        dy2, dx1 = b("bwd.dy2", T, BI, d.Hd), b("bwd.dx1", T, BI, d.Hd)
        dgi, dgh = b("bwd.dgi", T, BI, 3 * d.D), b("bwd.dgh", T, BI, 3 * d.D)
        dpin, dza = b("bwd.dpin", T, BI, d.Hd), b("bwd.dza", T, BI, d.Hd)
        dhp, dhc = b("bwd.dhp", T, BI, d.D), b("bwd.dhc", BI, d.D)
        dhin, dzin = b("bwd.dhin", T, BI, d.D), b("bwd.dzin", T, BI, d.Z)
        skinny = BI <= 128
        # ---- encoder backward (as a function of an image-row range)
        geo = ((64, 31, IC, cd), (31, 14, cd, 2 * cd), (14, 6, 2 * cd, 4 * cd), (6, 2, 4 * cd, 8 * cd))
        encgw = {}

        def enc_bwd_begin():
            for li in (1, 2, 3):
                _, _, ci, co = geo[li]
                if self.implicit_conv:
                    encgw[li] = b(f"bwd.gencwp{li}", co, 16 * ((ci + 31) // 32 * 32))   # channels padded to 32 per tap
                else:
                    encgw[li] = b(f"bwd.gencw{li}", co, 16 * ci)
                ops.fill(encgw[li], 0.0)

        def enc_bwd_rows(r0, r1):
            n = r1 - r0
            if I == 1:
                dea_c = dy2.view(N, d.Hd)[r0:r1]
            else:
                dea_c = b("bwd.dea_c", NB, d.Hd)[r0:r1]
                ops.group_sum(dy2.view(N, d.Hd)[r0 * I:r1 * I], I, dea_c)
            dembed = b("bwd.dembed", NB, d.E)[r0:r1]
            ops.gemm(dea_c, W(cell.post_mlp_e.weight), dembed, b_mn=True)
            da = b("bwd.da3", NB * 4, 8 * cd)[r0 * 4:r1 * 4]
            ops.permute4(dembed.view(n, 8 * cd, 4, 1), da.view(n, 4, 8 * cd, 1), (0, 2, 1, 3))
            for li in (3, 2, 1, 0):
                hin_, hout, ci, co = geo[li]
                hw = hout * hout
                if li == 3:                 # (layers 2..0: done by the col2im that produced their output gradient)
                    act = b(f"enc.a{li}", NB * hw, co)[r0 * hw:r1 * hw]
                    ops.bias_act_bwd(da, act, ACT_ELU, G(enc[2 * li].bias))
                if li == 0:
                    col = b(f"enc.col{li}", NB * hw, 16 * ci)[r0 * hw:r1 * hw]
                    ops.gemm(da, col, G(enc[0].weight).view(co, 16 * ci), a_mn=True, b_mn=True, accumulate=True)
                    continue
                if self.implicit_conv:
                    xprev = b(f"enc.a{li - 1}", NB * hin_ * hin_, ci)[r0 * hin_ * hin_:r1 * hin_ * hin_]
                    ops.conv_gemm(3, xprev.view(n, hin_, hin_, ci), 4, da, encgw[li])
                else:
                    col = b(f"enc.col{li}", NB * hw, 16 * ci)[r0 * hw:r1 * hw]
                    ops.gemm(da, col, encgw[li], a_mn=True, b_mn=True, accumulate=True)
                dcol = b(f"bwd.dcol{li}", NB * hw, 16 * ci)[r0 * hw:r1 * hw]
                ops.gemm(da, self._encw[li], dcol, b_mn=True)
                da_prev = b(f"bwd.da{li - 1}", NB * hin_ * hin_, ci)[r0 * hin_ * hin_:r1 * hin_ * hin_]
                act_prev = b(f"enc.a{li - 1}", NB * hin_ * hin_, ci)[r0 * hin_ * hin_:r1 * hin_ * hin_]
                # fold the column-form gradient back AND go through the ELU / bias of the layer below in the same pass
                ops.col2im_actbwd(dcol, hout, hout, 4, act_prev, G(enc[2 * (li - 1)].bias), da_prev.view(n, hin_, hin_, ci))
                da = da_prev

        def enc_bwd_end():
            for li in (1, 2, 3):
                _, _, ci, co = geo[li]
                if self.implicit_conv:
                    cpad = (ci + 31) // 32 * 32
                    ops.permute4(encgw[li].view(co, 4, 4, cpad)[..., :ci], G(enc[2 * li].weight), (0, 3, 1, 2))
                else:
                    ops.permute4(encgw[li].view(co, 4, 4, ci), G(enc[2 * li].weight), (0, 3, 1, 2))

        done = False
        if self._persistent_bptt_ok(BI):
            try:
                ops.rssm_unroll_bwd(
                    dict(T=T, BI=BI, D=d.D, Hd=d.Hd, G=d.G, C=d.C), conf.kl_weight, True,
                    ln2_g=self._raw(cell.post_norm.weight), ln1_g=self._raw(cell.in_norm.weight), post=post, pin=pin, y2=y2,
                    m2=m2, r2=r2, x1=x1, za=za, m1=m1, r1=r1, gates=gates, hin=hin, mask=mask, dfeat=dfeat3,
                    dpost_u=dpost_u, w=w3, dpost=dpost, dy2=dy2, dgi=dgi, dgh=dgh, dx1=dx1,
                    g_ln2_g=G(cell.post_norm.weight), g_ln2_b=G(cell.post_norm.bias), g_b_ph=G(cell.post_mlp_h.bias),
                    g_ln1_g=G(cell.in_norm.weight), g_ln1_b=G(cell.in_norm.bias), g_b_z=G(cell.z_mlp.bias),
                    ws_part2=b("k1b.part2", 4, BI, d.Hd), ws_part6=b("k1b.part6", 4, BI, d.D),
                    ws_part7=b("k1b.part7", 4, BI, d.Hd), ws_barrier=b("k1b.bar", 16, dtype=torch.int32), **self._k1b_w)
                done = True
            except RuntimeError as e:        # e.g. cooperative launch refused
                import warnings
                warnings.warn(f"pydreamer_b200: persistent BPTT kernel unavailable ({e}); using the per-timestep chain")
                self.persistent_bptt = False
        par = self._ov(2) and not done
        if skinny and not done:                     # the chain's split-K GEMMs reduce into pre-cleared outputs
            for buf_ in (dpin, dza, dhp, dhin, dzin):
                ops.fill(buf_, 0.0)
        for t in (() if done else reversed(range(T))):
            nxt = t < T - 1
            ops.cat_st_bwd(post[t], d.G, d.C, dfeat3[t, :, d.D:], dzin[t + 1] if nxt else None,
                           mask[t + 1] if nxt else None, dpost_u[t], w3[t], conf.kl_weight, dpost[t])
            ops.gemm(dpost[t], W(cell.post_mlp.weight), dpin[t], b_mn=True, c_zeroed=skinny)
            ops.ln_elu_bwd(dpin[t], y2[t], pin[t], self._raw(cell.post_norm.weight), m2[t], r2[t], dy2[t],
                           G(cell.post_norm.weight), G(cell.post_norm.bias), G(cell.post_mlp_h.bias))
            ops.gemm(dy2[t], W(cell.post_mlp_h.weight), dhp[t], b_mn=True, res=dfeat3[t, :, :d.D], c_zeroed=skinny)
            if par and nxt:
                self._join(2)                       # dhin[t + 1]
            ops.gru_bwd(dhp[t], dhin[t + 1] if nxt else None, mask[t + 1] if nxt else None, gates[t], hin[t], dgi[t],
                        dgh[t], dhc)
            if par:
                with self._fork(2):
                    ops.gemm(dgh[t], W(gru.weight_hh), dhin[t], b_mn=True, res=dhc, c_zeroed=skinny)
            else:
                ops.gemm(dgh[t], W(gru.weight_hh), dhin[t], b_mn=True, res=dhc, c_zeroed=skinny)
            ops.gemm(dgi[t], W(gru.weight_ih), dza[t], b_mn=True, c_zeroed=skinny)
            ops.ln_elu_bwd(dza[t], x1[t], za[t], self._raw(cell.in_norm.weight), m1[t], r1[t], dx1[t],
                           G(cell.in_norm.weight), G(cell.in_norm.bias), G(cell.z_mlp.bias))
            ops.gemm(dx1[t], W(cell.z_mlp.weight), dzin[t], b_mn=True, c_zeroed=skinny)
        if par:
            self._join(2)
        # batched weight gradients over all T*BI rows
        f2 = lambda x: x.view(N, x.shape[-1])
        ops.gemm(f2(dpost), f2(pin), G(cell.post_mlp.weight), a_mn=True, b_mn=True, accumulate=True)
        ops.colsum(f2(dpost), G(cell.post_mlp.bias))
        ops.gemm(f2(dy2), featN[:, :d.D], G(cell.post_mlp_h.weight), a_mn=True, b_mn=True, accumulate=True)
        ops.gemm(f2(dgh), f2(hin), G(gru.weight_hh), a_mn=True, b_mn=True, accumulate=True)
        ops.colsum(f2(dgh), G(gru.bias_hh))
        ops.gemm(f2(dgi), f2(za), G(gru.weight_ih), a_mn=True, b_mn=True, accumulate=True)
        ops.colsum(f2(dgi), G(gru.bias_ih))
        ops.gemm(f2(dx1), f2(zin), G(cell.z_mlp.weight), a_mn=True, b_mn=True, accumulate=True)
        if I == 1:
            dea, daa = f2(dy2), f2(dx1)
        else:
            dea, daa = b("bwd.dea", NB, d.Hd), b("bwd.daa", NB, d.Hd)
            ops.group_sum(f2(dy2), I, dea); ops.group_sum(f2(dx1), I, daa)
        embed = b("enc.embed", NB, d.E)
        ops.gemm(dea, embed, G(cell.post_mlp_e.weight), a_mn=True, b_mn=True, accumulate=True)
        ops.gemm(daa, obs["action"].reshape(NB, d.A), G(cell.a_mlp.weight), a_mn=True, b_mn=True, accumulate=True)
        enc_bwd_begin()
        enc_bwd_rows(0, NB)
        enc_bwd_end()
        if par_w:
            self._join(4)

    # ------------------------------------------------------------------ imagination rollout
    def _dream(self, feats, N, H, noise_actor, noise_prior, tag):
        """dreamer.py:188-216: H x { actor -> sample action -> forward_prior }, forward only (reinforce).
        feats (H+1, N, F): feats[0] holds the start states; rows 1..H are written here."""
        ops, d, conf = self.ops, self.d, self.conf
        b, W = (lambda name, *shape, **kw: self._buf(tag + name, *shape, **kw)), self._w
        cell = self.wm.core.cell
        gru = cell.gru.layers[0]
        ap = self._mlp_params(self.ac.actor)
        Ap = (d.Aout + 3) // 4 * 4                   # row pitch of the actor outputs: 16-byte rows keep them TMA-addressable
        alog = b("dream.alog", H, N, Ap)[..., :d.Aout]
        actions = b("dream.actions", H, N, d.A)
        aidx = b("dream.aidx", N, 1, dtype=torch.int32)
        aa, x, za = b("dream.aa", N, d.Hd), b("dream.x", N, d.Hd), b("dream.za", N, d.Hd)
        mm, rr = b("dream.m", N), b("dream.r", N)
        gi, gh = b("dream.gi", N, 3 * d.D), b("dream.gh", N, 3 * d.D)
        yp, pp, prior = b("dream.yp", N, d.Hd), b("dream.pp", N, d.Hd), b("dream.prior", N, d.Z)
        f16 = self.fp16_forward
        if f16:
            f16b = b("feats16", H + 1, N, d.F, dtype=torch.float16)
            ops.to_half(feats[0], f16b[0])
            za16, pp16 = b("dream.za16", N, d.Hd, dtype=torch.float16), b("dream.pp16", N, d.Hd, dtype=torch.float16)
            Wh = self._wh
        par = self._ov(2)
        if f16:
            gh_gemm = lambda i: ops.gemm_f16(f16b[i][:, :d.D], Wh(gru.weight_hh), gh, bias=self._raw(gru.bias_hh))
        else:
            gh_gemm = lambda i: ops.gemm(feats[i][:, :d.D], W(gru.weight_hh), gh, bias=self._raw(gru.bias_hh))
        if par:
            with self._fork(2):
                gh_gemm(0)
        for i in range(H):
            f, fn = feats[i], feats[i + 1]
            fh, fnh = (f16b[i], f16b[i + 1]) if f16 else (None, None)
            self._mlp_fwd(ap, f, alog[i], tag + "actor", rows_total=H * N, row0=i * N, save=True, x16=fh)
            if conf.actor_dist == "onehot":
                ops.cat_sample(alog[i], noise_actor[i], 1, d.A, actions[i], idx=aidx)
                if d.Hd % 4 == 0:
                    ops.gather_rows(aidx, self._waT, aa)             # a_mlp(one-hot) = one row of a_mlp^T
                else:
                    ops.gemm(actions[i], W(cell.a_mlp.weight), aa)
            else:
                ops.tanh_normal_sample(alog[i], noise_actor[i], actions[i])
                ops.gemm(actions[i], W(cell.a_mlp.weight), aa)
            if f16:
                ops.gemm_f16(fh[:, d.D:], Wh(cell.z_mlp.weight), x, bias=self._raw(cell.z_mlp.bias), res=aa)
                ops.ln_elu_fwd(x, self._raw(cell.in_norm.weight), self._raw(cell.in_norm.bias), 1e-3, za, mm, rr, za16)
                ops.gemm_f16(za16, Wh(gru.weight_ih), gi, bias=self._raw(gru.bias_ih))
                if par:
                    self._join(2)
                else:
                    gh_gemm(i)
                ops.gru_fwd(gi, gh, f[:, :d.D], fn[:, :d.D], h16=fnh[:, :d.D])
                if par and i + 1 < H:               # next step's h·W_hh overlaps prior MLP, sampling and the actor
                    with self._fork(2):
                        gh_gemm(i + 1)
                ops.gemm_f16(fnh[:, :d.D], Wh(cell.prior_mlp_h.weight), yp, bias=self._raw(cell.prior_mlp_h.bias))
                ops.ln_elu_fwd(yp, self._raw(cell.prior_norm.weight), self._raw(cell.prior_norm.bias), 1e-3, pp, mm, rr, pp16)
                ops.gemm_f16(pp16, Wh(cell.prior_mlp.weight), prior, bias=self._raw(cell.prior_mlp.bias))
                ops.cat_sample(prior, noise_prior[i], d.G, d.C, fn[:, d.D:], z16=fnh[:, d.D:])
            else:
                ops.gemm(f[:, d.D:], W(cell.z_mlp.weight), x, bias=self._raw(cell.z_mlp.bias), res=aa)
                ops.ln_elu_fwd(x, self._raw(cell.in_norm.weight), self._raw(cell.in_norm.bias), 1e-3, za, mm, rr)
                ops.gemm(za, W(gru.weight_ih), gi, bias=self._raw(gru.bias_ih))
                if par:
                    self._join(2)
                else:
                    gh_gemm(i)
                ops.gru_fwd(gi, gh, f[:, :d.D], fn[:, :d.D])
                if par and i + 1 < H:
                    with self._fork(2):
                        gh_gemm(i + 1)
                ops.gemm(fn[:, :d.D], W(cell.prior_mlp_h.weight), yp, bias=self._raw(cell.prior_mlp_h.bias))
                ops.ln_elu_fwd(yp, self._raw(cell.prior_norm.weight), self._raw(cell.prior_norm.bias), 1e-3, pp, mm, rr)
                ops.gemm(pp, W(cell.prior_mlp.weight), prior, bias=self._raw(cell.prior_mlp.bias))
                ops.cat_sample(prior, noise_prior[i], d.G, d.C, fn[:, d.D:])

    # ------------------------------------------------------------------ actor critic
    def _actor_critic(self, feats, N, H, want_grad, tag):
        """a2c.py:61-149 on the dreamed features (all inputs detached, dreamer.py:153-157)."""
        ops, d, conf, ac = self.ops, self.d, self.conf, self.ac
        J = H + 1
        b = lambda name, *shape, **kw: self._buf(tag + name, *shape, **kw)
        fall = feats.view(J * N, d.F)
        rp, tp = self._mlp_params(self.wm.decoder.reward.model), self._mlp_params(self.wm.decoder.terminal.model)
        cp, ctp, ap = self._mlp_params(ac.critic), self._mlp_params(ac.critic_target), self._mlp_params(ac.actor)
        rew, tlog = b("ac.rew", J * N, 1), b("ac.tlog", J * N, 1)
        vt, v = b("ac.vt", J * N, 1), b("ac.v", J * N, 1)
        fall16 = b("feats16", J, N, d.F, dtype=torch.float16).view(J * N, d.F) if self.fp16_forward else None
        self._mlp_fwd(rp, fall, rew, "scratch", x16=fall16)
        self._mlp_fwd(tp, fall, tlog, "scratch", x16=fall16)
        self._mlp_fwd(ctp, fall, vt, "scratch", x16=fall16)
        self._mlp_fwd(cp, fall, v, tag + "critic", save=True, x16=fall16)
        term = b("ac.term", J, N)
        adv, agae, target = b("ac.adv", H, N), b("ac.agae", H, N), b("ac.target", H, N)
        weight, dv = b("ac.weight", H, N), b("ac.dv", H * N, 1)
        sums = b("ac.sums", 8, dtype=torch.float64)
        ops.fill(sums.view(torch.float32), 0.0)
        ops.gae_critic(H, N, conf.gamma, conf.lambda_gae, vt, v, rew, tlog, term, adv, agae, target, weight, dv, sums)
        Ap = (d.Aout + 3) // 4 * 4
        alog = b("dream.alog", H, N, Ap).view(H * N, Ap)[:, :d.Aout]
        actions = b("dream.actions", H, N, d.A).view(H * N, d.A)
        dal = b("ac.dalog", H * N, Ap)[:, :d.Aout]
        if conf.actor_dist == "onehot":
            ops.actor_loss_onehot(conf.entropy, alog, actions, agae, weight, dal, sums[5:7])
        else:
            ops.actor_loss_tanh_normal(conf.entropy, alog, actions, agae, weight, dal, sums[5:7])
        if want_grad:
            fH = fall[:H * N]
            self._mlp_bwd(cp, fH, dv, tag + "critic", rows_total=J * N)
            self._mlp_bwd(ap, fH, dal, tag + "actor", rows_total=H * N)
        hm = float(H * N)
        s = sums
        r_mean = s[3] / hm
        r_var = torch.clamp((s[4] - s[3] * s[3] / hm) / (hm - 1.0), min=0.0)
        f32 = lambda x: x.to(torch.float32)
        metrics = dict(loss_critic=f32(s[0] / hm), loss_actor=f32(s[5] / hm), policy_entropy=f32(s[6] / hm),
                       policy_value=f32(s[1] / float(N)), policy_value_im=f32(s[2] / hm), policy_reward=f32(r_mean),
                       policy_reward_std=f32(r_var.sqrt()))
        return dict(loss_actor=metrics["loss_actor"], loss_critic=metrics["loss_critic"], metrics=metrics,
                    value=v.view(J, N), rew=rew.view(J, N), term=term,
                    tensors=dict(value=v.view(J, N), value_target=target, value_advantage=adv,
                                 value_advantage_gae=agae, value_weight=weight))

    def _dream_for_log(self, obs, T, B, I, noise_actor, noise_prior):
        """dreamer.py:165-180 (do_dream_tensors): dream T-1 steps from the first posterior state of every sequence, decode
        the imagined images, evaluate the critic (log_only).  Logging branch, no gradients."""
        ops, d = self.ops, self.d
        Hl, BI = T - 1, B * I
        N0 = T * B * I
        feats0 = self._buf("feats", self.imag_horizon + 1, N0, d.F) if ("feats", (self.imag_horizon + 1, N0, d.F), torch.float32) in self._ws \
            else next(v for k, v in self._ws.items() if k[0] == "feats" and k[1][1] == N0)
        fl = self._buf("dl.feats", Hl + 1, B, d.F)
        fl[0].copy_(feats0[0][0:BI:I])                                   # states[0, :, 0]
        self._dream(fl, B, Hl, noise_actor, noise_prior, "dl.")
        ac = self._actor_critic(fl, B, Hl, False, "dl.")
        img = self._image_decode(fl.view((Hl + 1) * B, d.F), (Hl + 1) * B, "dl.")
        actions = self._buf("dl.dream.actions", Hl, B, d.A)
        t = ac["tensors"]
        return dict(action_pred=torch.cat([obs["action"][:1], actions]), reward_pred=ac["rew"], terminal_pred=ac["term"],
                    image_pred=img.view(T, B, d.IC, 64, 64), value=t["value"], value_target=t["value_target"],
                    value_advantage=t["value_advantage"], value_advantage_gae=t["value_advantage_gae"],
                    value_weight=t["value_weight"])

    def _image_decode(self, featN, N, tag):
        """ConvDecoder.forward (decoders.py:157-161) without a loss: (N,F) -> (N,C,64,64)."""
        ops, d = self.ops, self.d
        cd, IC = d.cd, d.IC
        b = lambda name, *shape, **kw: self._buf(tag + name, *shape, **kw)
        dec = self.wm.decoder.image.model
        x0 = b("dec.x0", N, 32 * cd)
        ops.gemm(featN, self._w(dec[0].weight), x0, bias=self._raw(dec[0].bias), round_out=True)
        dgeo = ((1, 5, 5, 32 * cd, 4 * cd), (5, 13, 5, 4 * cd, 2 * cd), (13, 30, 6, 2 * cd, cd), (30, 64, 6, cd, IC))
        xin = x0
        for li, (hi, ho, k, ci, co) in enumerate(dgeo):
            cols = b(f"dec.cols{li}", N * hi * hi, k * k * co, dtype=self._cols_dtype(k * k * co))
            ops.gemm(xin, self._decw[li], cols)
            bias = self._raw(dec[2 + 2 * li].bias)
            if li < 3:
                a = b(f"dec.d{li}", N, ho, ho, co)
                ops.col2im(cols, hi, hi, k, bias, ACT_ELU, a, round_out=True)
                xin = a.view(N * ho * ho, co)
            else:
                out = b("dec.image", N, IC, 64, 64)
                ops.col2im(cols, hi, hi, k, bias, ACT_NONE, out.permute(0, 2, 3, 1), round_out=False)
        return out

    def __str__(self):
        n = sum(p.numel() for p in self.parameters())
        return f"Model: {n} parameters (pydreamer_b200: sm_100a kernels behind the pydreamer Dreamer API)"
