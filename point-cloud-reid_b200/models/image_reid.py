"""ImageReIDNet -- the token side of the reference's image re-identifier (mmdet3d/models/ReIDNet.py:839-1313): per-token
`downsample` (ReIDNet.py:1276-1277), `xcorr_eff` over two token sets with `cross_lin_attn` (ReIDNet.py:896-912),
`get_pooled_feats` (ReIDNet.py:1147-1155), `match_forward_inference` (ReIDNet.py:1057-1075), on the pcreid CUDA kernels.

The image backbone is a stock HuggingFace DeiT / BEiT (`get_image_model`, ReIDNet.py:811-834) and is NOT part of the
accelerated path: pass its token maps, or attach any module with `set_backbone`.  The matcher is
the same pair computation as the point-set `xcorr_eff` with the position code removed, so 'fast' mode reuses the fused
tcgen05 kernels (csrc/pair_tc2.cu) with a zero position image; token counts that are not a multiple of 128 (198 for
DeiT-distilled @224) run as zero-padded tiles.

New relative to the reference: ``match_all_pairs`` (dense (T, D) cost matrix without materialising per-pair gathers).
"""
import torch
from torch import nn

from .. import kernels as K
from .builder import FUSIONMODELS
from .lanegcn_nets import LinearRes
from .ReIDNet import _bias_cached, _cn, _i32, _kmajor_cached, build_module


def _seq_cn(seq, x, first_pm=False):
    """nn.Sequential of LinearRes / Linear on channel-major columns: x (B, C, R) -> (B, C', R).  first_pm: x is point-major
    (B, R, C) instead (only the first module reads it)."""
    for i, m in enumerate(seq):
        pm = first_pm and i == 0
        if isinstance(m, LinearRes):
            x = m.forward_cn(x, x_pm=pm)
        elif isinstance(m, nn.Linear):
            x = K.cn_linear(x, _kmajor_cached(m), bias=_bias_cached(m), x1_pm=pm)
        else:
            raise NotImplementedError(type(m))
    return x


@FUSIONMODELS.register_module()
class ImageReIDNet(nn.Module):
    def __init__(self, backbone, cls_head, match_head, vis_head, fp_head, downsample, cross_lin_attn, combine='cat', dim=768,
                 downsample_dim=128, losses_to_use=dict(kl=False, match=True, cls=True, shape=True, fp=True, triplet=True),
                 alpha=dict(kl=1, match=1, cls=1, shape=1, fp=1, triplet=1, vis=1), pool_type='both', compute_summary=True,
                 output_sequence_size=198, train_cfg=None, test_cfg=None, freeze_backbone=False, triplet_sample_num=5,
                 match_type='xcorr_eff', triplet_loss=dict(margin=0.2, p=2), eval_only=False):
        super().__init__()
        self.eval_only = eval_only
        if isinstance(backbone, nn.Module):
            self.backbone_name, self.backbone = type(backbone).__name__, backbone
        else:                       # 'deit-tiny', 'beit', ...: pretrained HuggingFace checkpoints, outside the accelerated path
            self.backbone_name, self.backbone = backbone, None
        self.image_processor = None
        self.cross_stage1 = build_module(dict(cross_lin_attn))
        self.cross_stage2 = build_module(dict(cross_lin_attn))
        self.cls_head = build_module(cls_head)
        self.match_head = build_module(match_head)
        self.vis_head = build_module(vis_head)
        self.fp_head = build_module(fp_head)
        self.downsample = build_module(downsample)
        self.combine = combine
        self.dim = dim
        self.downsample_dim = downsample_dim
        self.losses_to_use = dict(losses_to_use)
        self.compute_summary = compute_summary
        self.pool_type = pool_type
        self.output_sequence_size = output_sequence_size
        self.maxpool = nn.MaxPool1d(output_sequence_size)
        self.match_type = match_type
        self.bce = nn.BCEWithLogitsLoss()
        self.alpha = dict(alpha)
        self.verbose = False
        self.sampling = None
        self.match_mode = 'parity'
        self._fused = {}
        if self.match_type != 'xcorr_eff':
            raise NotImplementedError(f"match_type '{match_type}' is outside the accelerated hot path "
                                      "(every shipped image config uses 'xcorr_eff')")

    def set_mode(self, mode):
        """'parity': fp32 kernels (1e-4).  'parity_tc' / 'fast': fused tcgen05 matcher with fp16 / bf16 operands
        (|dlogit| <= 5e-3 / 3e-2)."""
        assert mode in ('parity', 'parity_tc', 'fast')
        self.match_mode = mode
        return self

    def fused_matcher(self):
        from . import fused_pairs
        fmt = fused_pairs.FMT_F16 if self.match_mode == 'parity_tc' else fused_pairs.FMT_BF16
        if fmt not in self._fused:
            self._fused[fmt] = fused_pairs.FusedXcorr(self, fmt)
        return self._fused[fmt]

    def set_backbone(self, module, name=None):
        """attach an image backbone (any module whose output has `.hidden_states` / `.last_hidden_state`, ReIDNet.py:914-941)."""
        self.backbone = module
        if name is not None:
            self.backbone_name = name
        return self

    # ------------------------------------------------------------------ token maps
    def _tokens(self, images):
        if self.backbone is None:
            raise RuntimeError(f"no image backbone attached ('{self.backbone_name}' is a pretrained HuggingFace model outside the "
                               "accelerated path): call set_backbone(module) or pass token maps to the *_tokens / match_* methods")
        with torch.no_grad():
            outputs = self.backbone(pixel_values=images)
        if 'deit' in self.backbone_name:
            return outputs.hidden_states[-1]
        if hasattr(outputs, 'last_hidden_state'):
            return outputs.last_hidden_state
        raise NotImplementedError("Not implemented for model: {}".format(self.backbone_name))

    def forward_inference(self, images):
        outputs = self._tokens(images)
        return self.get_pooled_feats(outputs.permute(0, 2, 1)), outputs

    def siamese_forward(self, sparse_1, sparse_2):
        assert sparse_1.shape == sparse_2.shape
        b = sparse_1.size(0)
        outputs = self._tokens(torch.cat([sparse_1, sparse_2], dim=0))
        return outputs[:b, ...].permute(0, 2, 1), outputs[b:, ...].permute(0, 2, 1)

    def downsample_tokens(self, h_cat):
        """h_cat (B, dim, S) -> (B, downsample_dim, S), exactly ``self.downsample(h_cat.reshape(-1,c)).reshape(b,dd,s)``
        (ReIDNet.py:1276-1277): the reference reshapes the channel-major buffer without permuting, so the rows the MLP sees
        are runs of `dim` consecutive floats; reproduced as written (rows in, rows out, no transposes)."""
        with torch.no_grad():
            b, c, s = h_cat.shape
            rows = h_cat.float().contiguous().reshape(1, b * s, c)                 # point-major rows of the flat buffer
            seq = list(self.downsample)
            x = _seq_cn(seq[:-1], rows, first_pm=True) if len(seq) > 1 else rows
            last = seq[-1]
            if not isinstance(last, nn.Linear):
                raise NotImplementedError("downsample is expected to end in a Linear (shipped image configs)")
            y = K.cn_linear(x, _kmajor_cached(last), bias=_bias_cached(last), x1_pm=len(seq) == 1, y_pm=True)   # (1, b*s, dd) rows
            return y.reshape(b, self.downsample_dim, s)

    # ------------------------------------------------------------------ pooling / heads
    def get_pooled_feats(self, h_cat):
        if self.pool_type == 'max':
            if h_cat.shape[1] != self.output_sequence_size:
                raise NotImplementedError("pool_type='max' with channels != output_sequence_size")
            return K.cn_chanmax(_cn(h_cat))
        if self.pool_type == 'both':
            return K.cn_pool(_cn(h_cat), mode=0)
        raise NotImplementedError

    def _head_cn(self, pooled_cn):
        """match_head on channel-major pooled features (1, C, P) -> logits (P,)."""
        return _seq_cn(self.match_head, pooled_cn).reshape(-1)

    def xcorr_eff(self, o1, o2, combine='add'):
        o1__ = self.cross_stage1(o1, o2)
        o2__ = self.cross_stage1(o2, o1)
        o1 = self.cross_stage2(o1__, o2__)
        o2 = self.cross_stage2(o2__, o1__)
        if self.combine == 'add':
            return o1 + o2
        if self.combine == 'minus':
            return o1 - o2
        if self.combine == 'cat':
            return torch.cat([o1, o2], dim=1)
        if self.combine == 'point-cat':
            return torch.cat([o1, o2], dim=2)
        raise NotImplementedError(self.combine)

    def match_forward_inference(self, h1, h2):
        """aligned pairs of (downsampled) token maps h1 / h2 (P, C, S) -> logits (P,)."""
        with torch.no_grad():
            ar = torch.arange(h1.shape[0], device=h1.device, dtype=torch.int32)
            return self._xcorr_pairs(_cn(h1), _cn(h2), ar, ar)

    # ------------------------------------------------------------------ all-pairs driver
    def _xcorr_pairs(self, h_t, h_d, ti, dj):
        """xcorr_eff + pooling + head for the pairs (ti[p], dj[p]) -> logits (P,), fp32 parity path; stage-1 queries and
        key/value summaries are per-object work, gathered per pair through index maps."""
        X1, X2 = self.cross_stage1, self.cross_stage2
        S_t, S_d = h_t.shape[2], h_d.shape[2]
        q_t, q_d = X1.search_query(h_t), X1.search_query(h_d)
        wkv_t, ks_t = X1.template_summary(h_t)
        wkv_d, ks_d = X1.template_summary(h_d)
        a = X1.attend(h_t, q_t, wkv_d, ks_d, S_d, s_map=ti, t_map=dj)
        b = X1.attend(h_d, q_d, wkv_t, ks_t, S_t, s_map=dj, t_map=ti)
        wkv_b, ks_b = X2.template_summary(b)
        wkv_a, ks_a = X2.template_summary(a)
        o1 = X2.attend(a, X2.search_query(a), wkv_b, ks_b, S_d)
        o2 = X2.attend(b, X2.search_query(b), wkv_a, ks_a, S_t)
        if self.pool_type == 'both' and self.combine == 'point-cat':
            return self._head_cn(K.cn_pool(o1, o2, mode=0, transposed=True))
        out = {'add': lambda: o1 + o2, 'minus': lambda: o1 - o2, 'cat': lambda: torch.cat([o1, o2], 1),
               'point-cat': lambda: torch.cat([o1, o2], 2)}[self.combine]()
        return self._head_cn(self.get_pooled_feats(out).t().contiguous().unsqueeze(0))

    def match_all_pairs(self, h_t, h_d, pair_mask=None, chunk=8192):
        """Dense (T, D) logit matrix over (downsampled) token maps h_t (T, C, S), h_d (D, C, S); entries where ``pair_mask``
        is False are 0."""
        with torch.no_grad():
            h_t, h_d = _cn(h_t), _cn(h_d)
            T, D = h_t.shape[0], h_d.shape[0]
            dev = h_t.device
            out = torch.zeros((T, D), device=dev, dtype=torch.float32)
            fused = None
            if self.match_mode in ('parity_tc', 'fast'):
                from . import fused_pairs
                if fused_pairs.supported(self, h_t.shape[2]) and h_t.shape[2] == h_d.shape[2]:
                    fused = self.fused_matcher()
                    pk_t, pk_d = fused.prepare(h_t, None), fused.prepare(h_d, None)
                    chunk = max(chunk, 65536 * 256 // h_t.shape[2])
            pairs = None if pair_mask is None else pair_mask.nonzero()
            total = T * D if pairs is None else pairs.shape[0]
            flat = out.view(-1)
            if pairs is None and fused is not None and D > 0:
                rows_per_chunk = max(1, chunk // D)
                for r0 in range(0, T, rows_per_chunk):
                    nrows = min(rows_per_chunk, T - r0)
                    flat[r0 * D:(r0 + nrows) * D] = fused.match(pk_t, pk_d, None, None, dense=(r0, nrows, D))
                return out
            for s in range(0, total, chunk):
                e = min(total, s + chunk)
                if pairs is None:
                    lin = torch.arange(s, e, device=dev)
                    ti, dj = lin // D, lin % D
                else:
                    ti, dj = pairs[s:e, 0], pairs[s:e, 1]
                    lin = ti * D + dj
                if fused is not None:
                    flat[lin] = fused.match(pk_t, pk_d, ti, dj)
                else:
                    flat[lin] = self._xcorr_pairs(h_t, h_d, _i32(ti), _i32(dj))
            return out

    # ------------------------------------------------------------------ mmdet BaseDetector surface
    def forward(self, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(**kwargs)
        return self.forward_test(**kwargs)

    def forward_train(self, *args, **kwargs):
        raise NotImplementedError("pcreid_b200 implements the inference hot path; train with the reference")

    def preprocess_inputs_size(self, sparse_1, sparse_2, vis_1, vis_2, label_1, label_2, id_1, id_2, size_1, size_2):
        st, ct = torch.stack, torch.cat
        return (st(sparse_1, 0), st(sparse_2, 0), ct(vis_1, 0), ct(vis_2, 0), ct(label_1, 0), ct(label_2, 0), ct(id_1, 0),
                ct(id_2, 0), ct(size_1, 0), ct(size_2, 0))

    def get_match_supervision(self, h1, h2, id_1, id_2):
        return h1, h2, (id_1 == id_2).float()

    def _aux_head(self, head, name, h):
        if head is None or not self.losses_to_use.get(name, False):
            return None
        pooled = self.get_pooled_feats(h).t().contiguous().unsqueeze(0)              # (1, C, B)
        return _seq_cn(head, pooled)[0].t().contiguous().squeeze(1)

    def forward_test(self, sparse_1, sparse_2, label_1, label_2, vis_1, vis_2, id_1, id_2, size_1, size_2, *args, **kwargs):
        """Result dict of ImageReIDNet.forward_test (ReIDNet.py:1248-1309): predictions of the match / cls / fp / vis heads;
        the kl and triplet evaluation losses (training diagnostics) are reported as 0."""
        (sparse_1, sparse_2, vis_1, vis_2, label_1, label_2, id_1, id_2, size_1, size_2) = \
            self.preprocess_inputs_size(sparse_1, sparse_2, vis_1, vis_2, label_1, label_2, id_1, id_2, size_1, size_2)
        h1, h2 = self.siamese_forward(sparse_1, sparse_2)
        h_cat = torch.cat([h1, h2], dim=0)
        labels, vis, ids = torch.cat([label_1, label_2], 0), torch.cat([vis_1, vis_2], 0), torch.cat([id_1, id_2], 0)
        with torch.no_grad():
            cls_preds = self._aux_head(self.cls_head, 'cls', h_cat)
            fp_preds = self._aux_head(self.fp_head, 'fp', h_cat)
            vis_filter = torch.where(torch.logical_and(ids != -1, vis != -1))
            vis_preds = self._aux_head(self.vis_head, 'vis', h_cat[vis_filter]) if vis_filter[0].numel() else None
        h1, h2, match = self.get_match_supervision(h1, h2, id_1, id_2)
        temp = self.downsample_tokens(h_cat)
        match_preds = self.match_forward_inference(temp[:h1.size(0)], temp[h1.size(0):])
        match_loss = self.bce(match_preds, match.to(match_preds.device)) * self.alpha['match']
        zero = torch.tensor([0.])
        results = {
            'val_fp_loss': zero, 'val_match_loss': torch.tensor([match_loss]), 'val_cls_loss': zero, 'val_kl_loss': zero,
            'val_triplet_loss': zero, 'val_vis_loss': zero, 'val_match_preds': match_preds, 'val_match_gt': match,
            'val_cls_preds': cls_preds, 'val_cls_gt': labels, 'val_vis_preds': vis_preds, 'val_vis_gt': vis[vis_filter],
            'val_fp_preds': fp_preds, 'val_fp_gt': (labels > 9).float(),
            'match_classes': torch.cat([label_1.unsqueeze(1), label_2.unsqueeze(1)], dim=1),
            'val_vis_gt_all': torch.cat([vis_1.unsqueeze(1), vis_2.unsqueeze(1)], dim=1),
            'num_points': torch.cat([size_1.unsqueeze(1), size_2.unsqueeze(1)], dim=1),
        }
        return [results]
