"""TEST INFRASTRUCTURE ONLY — functional torch-CPU restatement of the segmentation network.

Follows (reference): get_graph_feature[_with_normals] src/PointNet.py:72-140, DGCNNEncoderGn.forward :172-220,
PrimitivesEmbeddingDGCNGn.forward :265-289.  Weights come in as a state-dict with the reference's key names.
kNN uses the pinned-order C oracle (oracle/c/knn_oracle.c) unless `idx_list` is given.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .. import knn as oknn


def knn_idx(x_bcn, k, metric):
    x = x_bcn.detach().permute(0, 2, 1).contiguous().numpy()
    return torch.from_numpy(oknn.knn(x, k, metric))


def edge_features(x_bcn, idx):
    """(B,C,N),(B,N,k) -> (B,2C,N,k) = [x_j - x_i ; x_i]   (src/PointNet.py:86-102)"""
    B, C, N = x_bcn.shape
    k = idx.shape[-1]
    xt = x_bcn.permute(0, 2, 1)                                   # (B,N,C)
    nb = torch.gather(xt.unsqueeze(1).expand(B, N, N, C), 2, idx.unsqueeze(-1).expand(B, N, k, C))
    ctr = xt.unsqueeze(2).expand(B, N, k, C)
    return torch.cat([nb - ctr, ctr], 3).permute(0, 3, 1, 2)


def edge_conv_gn(x_bcn, idx, w, gn_w, gn_b, groups):
    f = edge_features(x_bcn, idx)
    y = F.conv2d(f, w)                                            # Conv2d 1x1, no bias (:157)
    y = F.leaky_relu(F.group_norm(y, groups, gn_w, gn_b), 0.2)
    return y.max(dim=-1)[0]


def encoder_fwd(sd, x, k, mode=5, idx_list=None, prefix="encoder."):
    """returns x4 (B,1024), x_features (B,256,N), [idx1, idx2, idx3]"""
    idxs = []
    cur = x
    outs = []
    for li, (cname, bname, groups) in enumerate([("conv1.0", "bn1", 2), ("conv2.0", "bn2", 2), ("conv3.0", "bn3", 2)]):
        if idx_list is not None:
            idx = idx_list[li]
        else:
            idx = knn_idx(cur, k, 1 if (mode == 5 and li == 0) else 0)
        idxs.append(idx)
        cur = edge_conv_gn(cur, idx, sd[prefix + cname + ".weight"], sd[prefix + bname + ".weight"],
                           sd[prefix + bname + ".bias"], groups)
        outs.append(cur)
    xf = torch.cat(outs, 1)
    y = F.conv1d(xf, sd[prefix + "mlp1.weight"], sd[prefix + "mlp1.bias"])
    y = F.relu(F.group_norm(y, 8, sd[prefix + "bnmlp1.weight"], sd[prefix + "bnmlp1.bias"]))
    return y.max(dim=2)[0], xf, idxs


def segnet_fwd(sd, points, k, mode=5, idx_list=None):
    """points (B,C,N) -> embedding (B,emb,N), primitives_log_prob (B,P,N)   (src/PointNet.py:265-284)"""
    B, _, N = points.shape
    x4, xf, idxs = encoder_fwd(sd, points, k, mode, idx_list)
    x = torch.cat([x4.view(B, 1024, 1).expand(B, 1024, N), xf], 1)
    x = F.relu(F.group_norm(F.conv1d(x, sd["conv1.weight"], sd["conv1.bias"]), 8, sd["bn1.weight"], sd["bn1.bias"]))
    x_all = F.relu(F.group_norm(F.conv1d(x, sd["conv2.weight"], sd["conv2.bias"]), 4, sd["bn2.weight"], sd["bn2.bias"]))
    e = F.relu(F.group_norm(F.conv1d(x_all, sd["mlp_seg_prob1.weight"], sd["mlp_seg_prob1.bias"]), 4,
                            sd["bn_seg_prob1.weight"], sd["bn_seg_prob1.bias"]))
    embedding = F.conv1d(e, sd["mlp_seg_prob2.weight"], sd["mlp_seg_prob2.bias"])
    p = F.relu(F.group_norm(F.conv1d(x_all, sd["mlp_prim_prob1.weight"], sd["mlp_prim_prob1.bias"]), 4,
                            sd["bn_prim_prob1.weight"], sd["bn_prim_prob1.bias"]))
    logits = F.conv1d(p, sd["mlp_prim_prob2.weight"], sd["mlp_prim_prob2.bias"])
    return embedding, F.log_softmax(logits, dim=1), idxs, x4, xf


def triplet_loss(output, labels, margin=1.0, rng=np.random):
    """EmbeddingLoss.triplet_loss (src/segment_loss.py:31-124) without mean-shift; host RNG draws are consumed in
    the reference's order (np.random.choice per label, then k1,k2 per iteration)."""
    B, D, N = output.shape
    out = F.normalize(output.permute(0, 2, 1), p=2, dim=2)
    total = out.new_zeros(1)
    only_one = 0
    samples = []
    for b in range(B):
        lab = labels[b]
        uniq = np.unique(lab)
        S = min(N // uniq.shape[0] + 1, 30)
        samples.append({l: rng.choice(list(np.where(lab == l)[0]), S, replace=True) for l in uniq})
    for b in range(B):
        keys = sorted(samples[b].keys())
        L = len(keys)
        if L == 1:
            only_one += 1
            continue
        shape_loss = out.new_zeros(1)
        norm = 0
        for _ in range(min(25, L * L)):
            k1 = rng.choice(L, 1)[0]
            k2 = rng.choice(L, 1)[0]
            if k1 == k2:
                continue
            norm += 1
            p1 = out[b, samples[b][keys[k1]]]
            p2 = out[b, samples[b][keys[k2]]]
            dp = ((p1[:, None] - p1[None]) ** 2).sum(2)
            dn = ((p1[:, None] - p2[None]) ** 2).sum(2)
            c = F.relu(dp - dn + margin)
            sat = ((c > 0).sum() + 1.0).float()
            shape_loss = shape_loss + (c.sum() - c.trace()) / sat.detach()
        total = total + shape_loss / (norm + 1e-8)
    return total / (B - only_one + 1e-8)
