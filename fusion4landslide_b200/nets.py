"""SURVEY 8(f) rank 2 -- the two small learned models that sit INSIDE the reference's per-patch loops, evaluated
for all patches of a tile at once (variable-length segments, CSR), parameter-compatible with the shipped weights:

    FilteringNetwork                 src/models/outlier_classifier.py:10-63   (12 x PointCN, 128 channels)
        per supervoxel in the reference (src/f2s3.py:340-347); here `compute_weights_segments` runs every
        supervoxel of a tile in one pass: the 1x1 convolutions are plain (K,128)x(128,128) products over all
        correspondences, InstanceNorm2d + BatchNorm2d(track_running_stats=False) of a single-sample batch are two
        per-segment, per-channel normalisations (biased variance, eps 1e-3).
    ClusterFeatureNetWithAttention   src/feature_aggregation/cluster_feature_net_self_attention.py:35-105
        per superpoint in the reference (`aggregation`, :72-103); `aggregate_segments` pools every superpoint of a
        tile at once (segment-wise softmax(QK^T/sqrt(d)) V -> fc -> mean -> MLP, centroid of the voxel coordinates).

The per-segment statistics and the segment-wise attention run in libf4l_b200.so (csrc/nets.cu); the dense
128x128 / 64x64 products are library GEMMs (torch.matmul -> cuBLAS), which is what a plain GEMM should be.
`state_dict` keys equal the reference modules', so `load_state_dict(torch.load('weights/...'))` works unchanged.
"""
import math

import torch
import torch.nn as nn

from . import ops

I32 = torch.int32


class PointCN(nn.Module):
    """Parameter container with the reference's layout (conv.0 / conv.4 are the two 1x1 convolutions)."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(channels, channels, kernel_size=1),
            nn.InstanceNorm2d(channels, eps=1e-3),
            nn.BatchNorm2d(channels, eps=1e-3, affine=False, track_running_stats=False),
            nn.ReLU(),
            nn.Conv2d(channels, channels, kernel_size=1),
            nn.InstanceNorm2d(channels, eps=1e-3),
            nn.BatchNorm2d(channels, eps=1e-3, affine=False, track_running_stats=False),
            nn.ReLU())

    def forward(self, x):
        return self.conv(x) + x


class FilteringNetwork(nn.Module):
    """Drop-in for `from src.models import FilteringNetwork` (main_f2s3.py:92-114)."""

    def __init__(self):
        super().__init__()
        numlayer, nchannel = 12, 128
        self.l1 = nn.Conv2d(6, nchannel, kernel_size=1)
        self.l2 = nn.Sequential(*[PointCN(nchannel) for _ in range(numlayer)])
        self.output = nn.Conv2d(nchannel, 1, kernel_size=1)
        self.activation = nn.ReLU(inplace=True)

    # -- the reference's single-supervoxel interface ---------------------------------------------------
    def compute_weights(self, x):
        """x (b,1,n,6) -> weights (b,n), outlier_classifier.py:52-63 (plain PyTorch forward, as in the reference)."""
        assert x.dim() == 4 and x.shape[1] == 1
        x = x.transpose(1, 3)
        # fp32 convolutions: cuDNN's default TF32 path changes the weights of this 24-times-normalised stack by up to
        # 0.2 (measured against the CPU fp32 forward with the shipped parameters), far beyond the 0.99999 gate
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            out = self.output(self.l2(self.l1(x))).squeeze(-1).squeeze(1)
        return self.activation(torch.tanh(out))

    def filter_input(self, data, data_raw, config):
        """outlier_classifier.py:65-105: weights from the network, then Kabsch -> res < coeff*median ->
        (>= 5 inliers and median < 0.5) -> refit, the tail in one launch (f4l_f2s3_prune_tail)."""
        from . import f2s3
        w = self.compute_weights(data)
        coeff = 2.5 if 'Rockfall_Simulator' in config.data_dir else 1
        n = data_raw.shape[1]
        ptr = torch.tensor([0, n], dtype=I32, device=data_raw.device)
        R, t, robust, _ = f2s3.filter_input_tail(data_raw[0], w[0], ptr, coeff)
        return {'scores': w, 'rot_est': R[0], 'trans_est': t[0].reshape(3, 1), 'robust_estimate': bool(robust[0].item())}

    # -- all supervoxels of a tile at once ----------------------------------------------------------------
    def compute_weights_segments(self, corr, seg_ptr, scale=True):
        """corr (K,6) rows grouped by supervoxel (CSR seg_ptr (Q+1) i32).  scale: divide every supervoxel's rows by
        its max |value| first (src/f2s3.py:343).  Returns the weights (K,) = relu(tanh(net))."""
        x = corr.to(torch.float32)
        ptr = seg_ptr.to(x.device, I32).contiguous()
        if scale:
            x = ops.segment_scale_maxabs(x.contiguous(), ptr)
        W = lambda conv: conv.weight.reshape(conv.weight.shape[0], -1).t().contiguous()
        h = torch.addmm(self.l1.bias, x, W(self.l1))
        for blk in self.l2:
            y = torch.addmm(blk.conv[0].bias, h, W(blk.conv[0]))
            y = ops.segment_norm2_relu(y, ptr, 1e-3)
            y = torch.addmm(blk.conv[4].bias, y, W(blk.conv[4]))
            h = ops.segment_norm2_relu(y, ptr, 1e-3, residual=h)
        out = torch.addmm(self.output.bias, h, W(self.output)).reshape(-1)
        return torch.relu(torch.tanh(out))


class SelfAttentionLayer(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim):
        super().__init__()
        self.query = nn.Linear(input_dim, hidden_dim)
        self.key = nn.Linear(input_dim, hidden_dim)
        self.value = nn.Linear(input_dim, hidden_dim)
        self.fc = nn.Linear(hidden_dim, output_dim)

    def forward(self, x):
        Q, K, V = self.query(x), self.key(x), self.value(x)
        a = torch.softmax(torch.matmul(Q, K.transpose(-2, -1)) / math.sqrt(K.size(-1)), dim=-1)
        return self.fc(torch.matmul(a, V))


class ClusterFeatureNetWithAttention(nn.Module):
    """Parameter-compatible with the reference module (weights/feat_aggregation_3d.pth, key 'state_dict')."""

    def __init__(self, cfg=None, input_feat_dim=64, hidden_feat_dim=64, output_feat_dim=64):
        super().__init__()
        if cfg is not None:
            input_feat_dim, hidden_feat_dim, output_feat_dim = cfg.input_feat_dim, cfg.hidden_feat_dim, cfg.output_feat_dim
        self.self_attention = SelfAttentionLayer(input_feat_dim, hidden_feat_dim, output_feat_dim)
        self.mlp = nn.Sequential(nn.Linear(output_feat_dim, hidden_feat_dim), nn.ReLU(),
                                 nn.Linear(hidden_feat_dim, output_feat_dim))

    def aggregate_segments(self, feats, coords, seg_ptr):
        """feats (V,D) / coords (V,3) of the voxels of every superpoint back to back (CSR seg_ptr (P+1) i32).
        Returns spt_feat (P,D_out), spt_coord (P,3) -- `aggregation` of the reference, mode 'test', for all
        superpoints at once.  mean_i fc(sum_j a_ij V_j) = fc(mean_i sum_j a_ij V_j): the fc layer is applied to
        the pooled row."""
        sa = self.self_attention
        ptr = seg_ptr.to(feats.device, I32).contiguous()
        x = feats.to(torch.float32)
        Q = torch.addmm(sa.query.bias, x, sa.query.weight.t())
        K = torch.addmm(sa.key.bias, x, sa.key.weight.t())
        V = torch.addmm(sa.value.bias, x, sa.value.weight.t())
        pooled = ops.segment_attention_pool(Q.contiguous(), K.contiguous(), V.contiguous(), ptr,
                                            1.0 / math.sqrt(K.shape[1]))       # (P, hidden) = mean_i softmax(.) V
        f = self.mlp(torch.addmm(sa.fc.bias, pooled, sa.fc.weight.t()))
        c = ops.segment_mean(coords.to(torch.float32).contiguous(), ptr)
        return f, c
