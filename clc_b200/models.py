"""Drop-in `CLC` and `TCM` models (reference: models/CLC_run.py:316-814, models/tcm.py:310-626).

Same constructor signature, submodule names (=> identical state_dict keys, checked by
tests/test_model_cpu.py against the reference itself) and `forward` contract
    forward(x, ref_frames=None) -> {"x_hat", "likelihoods": {"y", "z"}, "para": {"means", "scales", "y"}}
The analysis / synthesis / hyper transforms and the per-slice parameter networks are plain
PyTorch modules (clc_b200.layers); the latent hot path inside `forward` runs in the fused
sm_100a kernels:
    entropy_bottleneck(z) + z STE round  -> one launch   (CLC_run.py:526-530)
    per slice: gaussian_conditional + ste_round -> one launch (:569,:571); 0.5*tanh(lrp) add -> one (:582-583)
Opt-in extended wiring (`match_refs=True`, SURVEY.md 7.1 level C): each reference latent is first
aligned to y by Pearson patch matching + top-k gather (Patch_Matching.py) before entering
`ref_feature_adapter`.  The default (False) is the shipped reference forward.
"""
import math

import torch
import torch.nn as nn

from . import ops
from .entropy_models import EntropyBottleneck, GaussianConditional
from .layers import (CLMAlign, ConvTransBlock, ReferenceEncoder, ResidualBlockUpsample,
                     ResidualBlockWithStride, SWAtten, conv, conv1x1, conv3x3, subpel_conv3x3)
from .matching import match_and_gather

SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def _update_registered_buffers(module, module_name, buffer_names, state_dict, policy="resize_if_empty"):
    """Resize the (empty) CDF buffers so a checkpoint's tables can be loaded (CLC_run.py:46-97)."""
    valid = dict(module.named_buffers())
    for name in buffer_names:
        if name not in valid:
            raise ValueError(f'Invalid buffer name "{name}"')
    for name in buffer_names:
        key = f"{module_name}.{name}"
        if key not in state_dict:
            continue
        buf = valid[name]
        if policy in ("resize_if_empty", "resize"):
            if policy == "resize" or buf.numel() == 0:
                buf.resize_(state_dict[key].size())
        else:
            raise ValueError(f'Invalid policy "{policy}"')


class CompressionModel(nn.Module):
    def __init__(self, entropy_bottleneck_channels=None):
        super().__init__()
        if entropy_bottleneck_channels is not None:
            self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def update(self, force=False):
        updated = False
        for m in self.children():
            if isinstance(m, EntropyBottleneck):
                updated |= bool(m.update(force=force))
        return updated


def _stage(dim, head_dim, n, window=8):
    return [ConvTransBlock(dim, dim, head_dim, window, 0, "W" if i % 2 == 0 else "SW") for i in range(n)]


def _param_net(cin, cout=64):
    return nn.Sequential(conv(cin, 224, stride=1, kernel_size=3), nn.GELU(),
                         conv(224, 128, stride=1, kernel_size=3), nn.GELU(),
                         conv(128, cout, stride=1, kernel_size=3))


class _ChARMBase(CompressionModel):
    """Everything TCM and CLC share: transforms, slice attention, and the fused slice loop."""

    def _build_common(self, config, head_dim, drop_path_rate, N, M, num_slices, max_support_slices):
        if drop_path_rate:
            raise NotImplementedError("drop_path_rate > 0 is not used by the reference configs")
        self.config, self.head_dim, self.window_size = config, head_dim, 8
        self.num_slices, self.max_support_slices, self.M = num_slices, max_support_slices, M
        ws = self.window_size
        self.g_a = nn.Sequential(
            ResidualBlockWithStride(3, 2 * N, 2),
            *_stage(N, head_dim[0], config[0], ws), ResidualBlockWithStride(2 * N, 2 * N, stride=2),
            *_stage(N, head_dim[1], config[1], ws), ResidualBlockWithStride(2 * N, 2 * N, stride=2),
            *_stage(N, head_dim[2], config[2], ws), conv3x3(2 * N, M, stride=2))
        self.g_s = nn.Sequential(
            ResidualBlockUpsample(M, 2 * N, 2),
            *_stage(N, head_dim[3], config[3], ws), ResidualBlockUpsample(2 * N, 2 * N, 2),
            *_stage(N, head_dim[4], config[4], ws), ResidualBlockUpsample(2 * N, 2 * N, 2),
            *_stage(N, head_dim[5], config[5], ws), subpel_conv3x3(2 * N, 3, 2))

    def _build_hyper(self, config, N):
        self.h_a = nn.Sequential(ResidualBlockWithStride(320, 2 * N, 2), *_stage(N, 32, config[0], 4),
                                 conv3x3(2 * N, 192, stride=2))
        self.h_mean_s = nn.Sequential(ResidualBlockUpsample(192, 2 * N, 2), *_stage(N, 32, config[3], 4),
                                      subpel_conv3x3(2 * N, 320, 2))
        self.h_scale_s = nn.Sequential(ResidualBlockUpsample(192, 2 * N, 2), *_stage(N, 32, config[3], 4),
                                       subpel_conv3x3(2 * N, 320, 2))

    def _slice_width(self):
        return 320 // self.num_slices

    def _support_ch(self, i, extra=0):
        return 320 + self._slice_width() * min(i + extra, 5 + extra)

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated

    # -- fused latent path ---------------------------------------------------------------------
    def _hyper(self, y, noise):
        z = self.h_a(y)
        _, z_lik, z_hat = self.entropy_bottleneck(z, noise=None if noise is None else noise.get("z"),
                                                  ste=True, want_outputs=False)
        return z_lik, self.h_scale_s(z_hat), self.h_mean_s(z_hat)

    def _slice_loop(self, y, latent_means, latent_scales, ref_features, noise):
        """CLC_run.py:535-590 / tcm.py:440-475.  `transforms(i)` picks the per-slice networks."""
        y_shape = y.shape[2:]
        y_hat_slices, liks, mus, scales = [], [], [], []
        use_ref = ref_features is not None
        # likelihoods and y_hat of every slice are written in place into [B, 320, h, w] buffers by the
        # per-slice launches (batch-strided kernel arguments): no torch.cat of them (CLC_run.py:587,:590)
        inplace = getattr(self, "_inplace_slices", True)     # (the CPU oracle adapters of the tests turn it off)
        bufs = ops.SliceBuffers(y.shape[0], y.shape[1], y_shape[0], y_shape[1], y.device) if inplace else None
        c0 = 0
        # the mean and the scale branch of a slice (SWAtten + 3 convolutions each, CLC_run.py:537-566) are
        # independent: with `fork_branches` the scale branch runs on a side stream -- two parallel branches once
        # the forward is captured into a CUDA graph (make_graphed_forward)
        fork = getattr(self, "fork_branches", False) and y.is_cuda and not torch.is_grad_enabled()
        for i, y_slice in enumerate(y.chunk(self.num_slices, 1)):
            support = y_hat_slices if self.max_support_slices < 0 else y_hat_slices[:self.max_support_slices]

            def scale_branch():
                ss = self.atten_scale[i](torch.cat([latent_scales] + support, dim=1))
                if use_ref:
                    return self.ref_cc_scale_transforms[i](torch.cat([ss, ref_features], dim=1))
                return self.cc_scale_transforms[i](ss)

            if fork:
                cur, side = torch.cuda.current_stream(y.device), self._side_stream(y.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    scale = scale_branch()
                    scale.record_stream(cur)
            mean_support = self.atten_mean[i](torch.cat([latent_means] + support, dim=1))
            if use_ref:
                mu = self.ref_cc_mean_transforms[i](torch.cat([mean_support, ref_features], dim=1))
            else:
                mu = self.cc_mean_transforms[i](mean_support)
            if fork:
                cur.wait_stream(side)
            else:
                scale = scale_branch()
            mu = mu[:, :, :y_shape[0], :y_shape[1]]
            scale = scale[:, :, :y_shape[0], :y_shape[1]]
            mus.append(mu)
            scales.append(scale)
            # one launch: likelihood (train: y + U(-.5,.5); eval: round) AND ste_round(y - mu) + mu
            n_i = None if noise is None else noise["y"][:, i * y_slice.shape[1]:(i + 1) * y_slice.shape[1]]
            if inplace:
                _, lik_i, y_hat_i = self.gaussian_conditional(y_slice, scale, mu, noise=n_i, ste=True,
                                                              want_outputs=False,
                                                              out=bufs.take(c0, c0 + y_slice.shape[1]))
            else:
                _, lik_i, y_hat_i = self.gaussian_conditional(y_slice, scale, mu, noise=n_i, ste=True,
                                                              want_outputs=False)
            c0 += y_slice.shape[1]
            liks.append(lik_i)
            if use_ref:
                lrp = self.ref_lrp_transforms[i](torch.cat([mean_support, y_hat_i, ref_features], dim=1))
            else:
                lrp = self.lrp_transforms[i](torch.cat([mean_support, y_hat_i], dim=1))
            y_hat_i = self._lrp_add(y_hat_i, lrp)  # y_hat += 0.5 * tanh(lrp), in place
            y_hat_slices.append(y_hat_i)
        if not inplace:
            return (torch.cat(y_hat_slices, dim=1), torch.cat(mus, dim=1), torch.cat(scales, dim=1),
                    torch.cat(liks, dim=1))
        return (ops.assemble(bufs, "y_hat", y_hat_slices), torch.cat(mus, dim=1), torch.cat(scales, dim=1),
                ops.assemble(bufs, "lik", liks))

    @staticmethod
    def _lrp_add(y_hat, lrp):
        return ops.lrp_add_(y_hat, lrp)

    def _side_stream(self, device):
        st = getattr(self, "_side", None)
        if st is None or st.device != device:
            st = self._side = torch.cuda.Stream(device=device)
        return st

    def make_graphed_forward(self, *example_inputs, fork_branches=True, channels_last=False, tf32_matmul=False,
                             warmup=3):
        """Inference: capture `forward` for these input shapes into ONE CUDA graph and return `run(*inputs) ->
        the same output dict` (tensors are static buffers, overwritten by the next call).  The reference's
        forward is ~1 500 eager launches (SWAtten blocks, the 5-slice ChARM loop with 10 parameter networks),
        launch-bound at every image size it is evaluated on; the graph removes the host from the loop and, with
        `fork_branches`, runs each slice's mean and scale branch in parallel.  `channels_last=True` additionally
        converts the weights (in place) and the static inputs to torch.channels_last: cuDNN then runs its NHWC
        kernels without the per-call NCHW<->NHWC conversion kernels (1.4 ms of 12 ms at 256 x 256) and the Swin
        blocks' 'b c h w -> b h w c' rearranges become views; results then differ from the NCHW forward by
        convolution round-off (other algorithms), not bit for bit.  `tf32_matmul=True` lets the Linear layers of the
        Swin blocks use TF32 tensor cores while the graph is captured (the convolutions already do, cuDNN's default,
        as in the reference; the reference's matmuls are fp32, so this is a stated deviation: ~1e-3 relative in the
        activations)."""
        if self.training:
            raise RuntimeError("make_graphed_forward is for eval mode (training draws fresh noise per step: use "
                               "clc_b200.latent_path.LatentPath for a captured training step)")

        if channels_last:
            self.to(memory_format=torch.channels_last)

        def clone(a):
            if isinstance(a, torch.Tensor):
                if channels_last and a.dim() == 4:
                    return a.detach().clone(memory_format=torch.channels_last)
                return a.detach().clone()
            if isinstance(a, (list, tuple)):
                return [clone(t) for t in a]
            return a

        def copy_in(dst, src):
            if isinstance(dst, torch.Tensor):
                dst.copy_(src)
            elif isinstance(dst, list):
                if len(dst) != len(src):
                    raise ValueError("graphed forward: number of reference frames differs from the captured call")
                for d, t in zip(dst, src):
                    copy_in(d, t)

        static = [clone(a) for a in example_inputs]
        dev = next(self.parameters()).device
        prev, self.fork_branches = getattr(self, "fork_branches", False), bool(fork_branches)
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32_matmul) or prev_tf32
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(warmup):
                    self(*static)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                out = self(*static)
        finally:
            self.fork_branches = prev
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32

        def run(*inputs):
            if len(inputs) != len(static):
                raise ValueError("graphed forward: call with the same arguments as the captured example")
            with torch.no_grad():
                for d, t in zip(static, inputs):
                    copy_in(d, t)
            graph.replay()
            return out

        run.graph = graph
        return run

    # -- bitstreams (CLC_run.py:629-716, :738-814; tcm.py compress / decompress) ---------------------
    def _slice_params(self, i, latent_means, latent_scales, y_hat_slices, ref_features, y_shape):
        support = y_hat_slices if self.max_support_slices < 0 else y_hat_slices[:self.max_support_slices]
        mean_support = self.atten_mean[i](torch.cat([latent_means] + support, dim=1))
        scale_support = self.atten_scale[i](torch.cat([latent_scales] + support, dim=1))
        if ref_features is not None:
            mu = self.ref_cc_mean_transforms[i](torch.cat([mean_support, ref_features], dim=1))
            scale = self.ref_cc_scale_transforms[i](torch.cat([scale_support, ref_features], dim=1))
        else:
            mu = self.cc_mean_transforms[i](mean_support)
            scale = self.cc_scale_transforms[i](scale_support)
        return mean_support, mu[:, :, :y_shape[0], :y_shape[1]], scale[:, :, :y_shape[0], :y_shape[1]]

    def _slice_refine(self, i, mean_support, y_hat_i, ref_features):
        if ref_features is not None:
            lrp = self.ref_lrp_transforms[i](torch.cat([mean_support, y_hat_i, ref_features], dim=1))
        else:
            lrp = self.lrp_transforms[i](torch.cat([mean_support, y_hat_i], dim=1))
        return self._lrp_add(y_hat_i, lrp)

    def _compress(self, x, ref_features):
        """Same stream layout as the reference: one string per image for z (EntropyBottleneck.compress),
        ONE string for y holding, slice after slice, the symbols of the whole batch (CLC_run.py:693,:712).
        Symbols and scale-table indexes are produced on the device by clc_gc_symbols_indexes and reach
        the host coder in a single copy (the reference: two `.tolist()` per slice)."""
        from .ans import BufferedRansEncoder
        gc = self.gaussian_conditional
        tables = gc.coder_tables()
        y = self.g_a(x)
        y_shape = y.shape[2:]
        z = self.h_a(y)
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        latent_scales, latent_means = self.h_scale_s(z_hat), self.h_mean_s(z_hat)
        table = gc.scale_table.to(x.device)
        y_hat_slices, symbols, indexes = [], [], []
        for i, y_slice in enumerate(y.chunk(self.num_slices, 1)):
            mean_support, mu, scale = self._slice_params(i, latent_means, latent_scales, y_hat_slices,
                                                         ref_features, y_shape)
            sym, idx = ops.gc_symbols_indexes(y_slice, scale, mu, table)
            symbols.append(sym.reshape(-1))
            indexes.append(idx.reshape(-1))
            y_hat_slices.append(self._slice_refine(i, mean_support, sym.to(torch.float32) + mu, ref_features))
        encoder = BufferedRansEncoder()
        encoder.encode_with_indexes(torch.cat(symbols), torch.cat(indexes), tables, None, None)
        return {"strings": [[encoder.flush()], z_strings], "shape": z.size()[-2:]}

    def _decompress(self, strings, shape, ref_features):
        from .ans import RansDecoder
        gc = self.gaussian_conditional
        tables = gc.coder_tables()
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        latent_scales, latent_means = self.h_scale_s(z_hat), self.h_mean_s(z_hat)
        y_shape = [z_hat.shape[2] * 4, z_hat.shape[3] * 4]
        decoder = RansDecoder()
        decoder.set_stream(strings[0][0])
        y_hat_slices = []
        for i in range(self.num_slices):
            mean_support, mu, scale = self._slice_params(i, latent_means, latent_scales, y_hat_slices,
                                                         ref_features, y_shape)
            index = gc.build_indexes(scale)                                    # device kernel
            rv = decoder.decode_stream(index.reshape(-1), tables, None, None, as_tensor=True)
            y_hat_i = gc.dequantize(rv.reshape(mu.shape).to(mu.device), mu)
            y_hat_slices.append(self._slice_refine(i, mean_support, y_hat_i, ref_features))
        y_hat = torch.cat(y_hat_slices, dim=1)
        return {"x_hat": self.g_s(y_hat).clamp_(0, 1)}


class TCM(_ChARMBase):
    def __init__(self, config=[2, 2, 2, 2, 2, 2], head_dim=[8, 16, 32, 32, 16, 8], drop_path_rate=0, N=128, M=320,
                 num_slices=5, max_support_slices=5, **kwargs):
        super().__init__(entropy_bottleneck_channels=N)
        self._build_common(config, head_dim, drop_path_rate, N, M, num_slices, max_support_slices)
        self._build_hyper(config, N)
        ws, sw = self.window_size, self._slice_width()
        self.atten_mean = nn.ModuleList(nn.Sequential(SWAtten(self._support_ch(i), self._support_ch(i), 16, ws, 0, inter_dim=128))
                                        for i in range(num_slices))
        self.atten_scale = nn.ModuleList(nn.Sequential(SWAtten(self._support_ch(i), self._support_ch(i), 16, ws, 0, inter_dim=128))
                                         for i in range(num_slices))
        self.cc_mean_transforms = nn.ModuleList(_param_net(self._support_ch(i), sw) for i in range(num_slices))
        self.cc_scale_transforms = nn.ModuleList(_param_net(self._support_ch(i), sw) for i in range(num_slices))
        self.lrp_transforms = nn.ModuleList(_param_net(self._support_ch(i, 1), sw) for i in range(num_slices))
        self.entropy_bottleneck = EntropyBottleneck(192)
        self.gaussian_conditional = GaussianConditional(None)

    def forward(self, x, noise=None):
        y = self.g_a(x)
        z_lik, latent_scales, latent_means = self._hyper(y, noise)
        y_hat, means, scales, y_lik = self._slice_loop(y, latent_means, latent_scales, None, noise)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik},
                "para": {"means": means, "scales": scales, "y": y}}

    def compress(self, x):
        """tcm.py compress: {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}."""
        return self._compress(x, None)

    def decompress(self, strings, shape):
        return self._decompress(strings, shape, None)

    def load_state_dict(self, state_dict, strict=True):
        _update_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                   ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        _update_registered_buffers(self.entropy_bottleneck, "entropy_bottleneck",
                                   ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
        return super().load_state_dict(state_dict, strict=strict)

    @classmethod
    def from_state_dict(cls, state_dict):
        N = state_dict["g_a.0.weight"].size(0)
        M = state_dict["g_a.6.weight"].size(0)
        net = cls(N, M)
        net.load_state_dict(state_dict)
        return net


class CLC(_ChARMBase):
    def __init__(self, config=[2, 2, 2, 2, 2, 2], head_dim=[8, 16, 32, 32, 16, 8], drop_path_rate=0, N=128, M=320,
                 num_slices=5, max_support_slices=5, num_ref_frames=3, use_ref=True, match_refs=False,
                 match_patch=4, match_k=4, match_temperature=15.0, match_mode="tc", **kwargs):
        super().__init__(entropy_bottleneck_channels=N)
        self.num_ref_frames, self.use_ref = num_ref_frames, use_ref
        self.match_refs, self.match_patch, self.match_k = match_refs, match_patch, match_k
        self.match_temperature, self.match_mode = match_temperature, match_mode
        self._build_common(config, head_dim, drop_path_rate, N, M, num_slices, max_support_slices)
        self.ref_encoder = ReferenceEncoder(N, M)
        # Instantiated-but-unused in the reference forward (CLC_run.py:359-369); kept for key parity.
        self.feature_alignment = nn.ModuleList(CLMAlign(192, head_dim=32, window_size=4) for _ in range(num_ref_frames))
        self.multi_ref_fusion = nn.Sequential(conv1x1(192 * (num_ref_frames + 1), 256), nn.GELU(), conv1x1(256, 192))
        self._build_hyper(config, N)
        ws, sw = self.window_size, self._slice_width()
        self.atten_mean = nn.ModuleList(nn.Sequential(SWAtten(self._support_ch(i), self._support_ch(i), 16, ws, 0, inter_dim=128))
                                        for i in range(num_slices))
        self.atten_scale = nn.ModuleList(nn.Sequential(SWAtten(self._support_ch(i), self._support_ch(i), 16, ws, 0, inter_dim=128))
                                         for i in range(num_slices))
        self.ref_cc_mean_transforms = nn.ModuleList(_param_net(self._support_ch(i) + 64, sw) for i in range(num_slices))
        self.ref_cc_scale_transforms = nn.ModuleList(_param_net(self._support_ch(i) + 64, sw) for i in range(num_slices))
        self.cc_mean_transforms = nn.ModuleList(_param_net(self._support_ch(i), sw) for i in range(num_slices))
        self.cc_scale_transforms = nn.ModuleList(_param_net(self._support_ch(i), sw) for i in range(num_slices))
        self.lrp_transforms = nn.ModuleList(_param_net(self._support_ch(i, 1), sw) for i in range(num_slices))
        self.ref_lrp_transforms = nn.ModuleList(_param_net(self._support_ch(i, 1) + 64, sw) for i in range(num_slices))
        self.ref_feature_adapter = nn.Sequential(conv1x1(M * num_ref_frames, 128), nn.GELU(), conv1x1(128, 64))
        self.entropy_bottleneck = EntropyBottleneck(192)
        self.gaussian_conditional = GaussianConditional(None)

    def extract_ref_features(self, ref_frames, y=None):
        """CLC_run.py:493-510; with match_refs the reference latents are aligned to y first."""
        if ref_frames is None or not self.use_ref:
            return None
        B = ref_frames[0].size(0)
        # one encoder pass over all references instead of a Python loop of n_refs passes
        feats = self.ref_encoder(torch.cat(list(ref_frames), dim=0))
        R = len(ref_frames)
        feats = feats.view(R, B, *feats.shape[1:]).transpose(0, 1)  # [B, R, M, h, w]
        if self.match_refs:
            if y is None:
                raise ValueError("match_refs needs the image latent y")
            feats = self._align_refs(y, feats.contiguous())
        return self.ref_feature_adapter(feats.reshape(B, R * feats.shape[2], *feats.shape[3:]))

    def _align_refs(self, y, feats):
        """Level-C wiring: every reference latent is replaced by its patch-matched, softmax-blended
        version, SI_Finder_at_Decoder_Feature_Domain(y, ref, ...)['1'] (Patch_Matching.py:157-216) with
        ph = pw = match_patch, num_k = match_k, temperature, Gaussian mask on.  feats [B, R, M, h, w]."""
        return match_and_gather(y, feats, self.match_patch, self.match_patch, self.match_k,
                                self.match_temperature, True, False, self.match_mode)

    def forward(self, x, ref_frames=None, noise=None):
        """`noise` (optional dict {"y": [B,320,h,w], "z": [B,192,h/4,w/4]}) injects the training
        U(-1/2,1/2) samples explicitly for bit-reproducible parity runs."""
        y = self.g_a(x)
        ref_features = self.extract_ref_features(ref_frames, y if self.match_refs else None)
        z_lik, latent_scales, latent_means = self._hyper(y, noise)
        y_hat, means, scales, y_lik = self._slice_loop(y, latent_means, latent_scales,
                                                       ref_features if self.use_ref else None, noise)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik},
                "para": {"means": means, "scales": scales, "y": y}}

    def _coder_ref_features(self, ref_frames):
        if self.match_refs:
            raise NotImplementedError("match_refs aligns the references to y, which a decoder does not have: "
                                      "bitstreams are defined for the shipped wiring (match_refs=False)")
        return self.extract_ref_features(ref_frames) if self.use_ref else None

    def compress(self, x, ref_frames=None):
        """CLC_run.py:629-716 -> {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}.
        Call update() first (as with the reference) so the CDF tables exist."""
        return self._compress(x, self._coder_ref_features(ref_frames))

    def decompress(self, strings, shape, ref_frames=None):
        """CLC_run.py:738-814 -> {"x_hat"} clamped to [0, 1]."""
        return self._decompress(strings, shape, self._coder_ref_features(ref_frames))

    def symbols_and_indexes(self, x, ref_frames=None):
        """Device-resident coder inputs of `compress` (CLC_run.py:629-716): per-slice int32 symbols
        (:690) and scale-table indexes (:689), without the reference's `.tolist()` host copies."""
        if self.gaussian_conditional.scale_table.numel() == 0:
            self.update()
        y = self.g_a(x)
        ref_features = self.extract_ref_features(ref_frames, y if self.match_refs else None)
        z = self.h_a(y)
        _, _, z_hat = self.entropy_bottleneck(z, training=False, ste=True, want_outputs=False)
        latent_scales, latent_means = self.h_scale_s(z_hat), self.h_mean_s(z_hat)
        y_shape = y.shape[2:]
        y_hat_slices, symbols, indexes = [], [], []
        table = self.gaussian_conditional.scale_table.to(x.device)
        for i, y_slice in enumerate(y.chunk(self.num_slices, 1)):
            support = y_hat_slices if self.max_support_slices < 0 else y_hat_slices[:self.max_support_slices]
            mean_support = self.atten_mean[i](torch.cat([latent_means] + support, dim=1))
            scale_support = self.atten_scale[i](torch.cat([latent_scales] + support, dim=1))
            if ref_features is not None:
                mu = self.ref_cc_mean_transforms[i](torch.cat([mean_support, ref_features], dim=1))
                scale = self.ref_cc_scale_transforms[i](torch.cat([scale_support, ref_features], dim=1))
            else:
                mu = self.cc_mean_transforms[i](mean_support)
                scale = self.cc_scale_transforms[i](scale_support)
            mu = mu[:, :, :y_shape[0], :y_shape[1]]
            scale = scale[:, :, :y_shape[0], :y_shape[1]]
            sym, idx = ops.gc_symbols_indexes(y_slice, scale, mu, table)
            symbols.append(sym)
            indexes.append(idx)
            y_hat_i = sym.to(torch.float32) + mu
            if ref_features is not None:
                lrp = self.ref_lrp_transforms[i](torch.cat([mean_support, y_hat_i, ref_features], dim=1))
            else:
                lrp = self.lrp_transforms[i](torch.cat([mean_support, y_hat_i], dim=1))
            y_hat_slices.append(self._lrp_add(y_hat_i, lrp))
        return {"symbols": torch.cat(symbols, 1), "indexes": torch.cat(indexes, 1),
                "y_hat": torch.cat(y_hat_slices, 1)}

    def load_state_dict(self, state_dict, strict=False):
        """Non-strict, filtered load keeping compatibility with TCM / older checkpoints
        (CLC_run.py:599-618)."""
        own = self.state_dict()
        filtered = {k: v for k, v in state_dict.items() if k in own}
        _update_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                   ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        # checkpoints saved after update() also carry the EntropyBottleneck tables (compressai's
        # CompressionModel.load_state_dict resizes the buffers of every entropy model)
        _update_registered_buffers(self.entropy_bottleneck, "entropy_bottleneck",
                                   ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
        return super().load_state_dict(filtered, strict=False)

    @classmethod
    def from_state_dict(cls, state_dict):
        N = state_dict["g_a.0.weight"].size(0)
        M = state_dict["g_a.6.weight"].size(0)
        net = cls(N=N // 2, M=M)
        net.load_state_dict(state_dict)
        return net
