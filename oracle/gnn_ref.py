"""oracle/gnn_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py; parity unpinned).

Plain-PyTorch (CPU, fp32, autograd) restatement of the torch_geometric layers the
reference's Q-network calls, and of the two networks themselves:

* ``SAGEConv`` / ``GCNConv`` / ``TopKPooling`` / ``global_max_pool`` /
  ``global_mean_pool``  -- PyG 2.0-2.2 semantics, SURVEY.md Appendix A.10
  (call sites: /root/reference/airfoilgcnn.py:30-41,94-122)
* ``NodeRemovalNet``    -- /root/reference/airfoilgcnn.py:24-145
* ``AirfoilGCNN``       -- /root/reference/airfoilgcnn.py:148-209
* ``replay_loss``       -- /root/reference/airfoil_dqn.py:240-310

torch_geometric is not vendored under /root/reference and is absent from this
image; state_dict key names follow PyG so reference checkpoints would load.
TopK ties are pinned to "descending score, ties -> lower node index".
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


def scatter_sum(src, index, n):
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, index, src)


class _Lin(nn.Module):
    def __init__(self, i, o, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))
        self.bias = nn.Parameter(torch.empty(o)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            bound = 1 / math.sqrt(i)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class SAGEConv(nn.Module):
    """lin_l(mean_{j->i} x_j) + lin_r(x_i); directed edges as given, duplicates counted."""

    def __init__(self, i, o):
        super().__init__()
        self.lin_l = _Lin(i, o, bias=True)
        self.lin_r = _Lin(i, o, bias=False)

    def forward(self, x, edge_index):
        n = x.shape[0]
        src, dst = edge_index[0], edge_index[1]
        s = scatter_sum(x[src], dst, n)
        cnt = scatter_sum(torch.ones(len(dst), dtype=x.dtype), dst, n).clamp(min=1)
        return self.lin_l(s / cnt[:, None]) + self.lin_r(x)


class GCNConv(nn.Module):
    """D^-1/2 (A+I) D^-1/2 (x W) + b with self loops replaced by one weight-1 loop per node."""

    def __init__(self, i, o):
        super().__init__()
        self.lin = _Lin(i, o, bias=False)
        self.bias = nn.Parameter(torch.zeros(o))

    def forward(self, x, edge_index):
        n = x.shape[0]
        src, dst = edge_index[0], edge_index[1]
        keep = src != dst
        loop = torch.arange(n, dtype=src.dtype)
        src = torch.cat([src[keep], loop])
        dst = torch.cat([dst[keep], loop])
        w = torch.ones(len(src), dtype=x.dtype)
        deg = scatter_sum(w, dst, n)
        dis = deg.pow(-0.5)
        dis[torch.isinf(dis)] = 0
        norm = dis[src] * w * dis[dst]
        xw = self.lin(x)
        return scatter_sum(norm[:, None] * xw[src], dst, n) + self.bias


def topk_perm(score, ratio, batch, num_graphs):
    """Per graph keep ceil(ratio*n) (float32 arithmetic as PyG) largest; ties -> lower index."""
    n_per = scatter_sum(torch.ones_like(batch), batch, num_graphs)
    k_per = (ratio * n_per.to(torch.float)).ceil().to(torch.long)
    perms = []
    start = 0
    for g in range(num_graphs):
        n = int(n_per[g])
        s = score[start:start + n]
        order = torch.sort(s, descending=True, stable=True).indices
        perms.append(order[: int(k_per[g])] + start)
        start += n
    return torch.cat(perms) if perms else torch.zeros(0, dtype=torch.long)


class TopKPooling(nn.Module):
    def __init__(self, c, ratio):
        super().__init__()
        self.ratio = ratio
        self.weight = nn.Parameter(torch.empty(1, c))
        bound = 1.0 / math.sqrt(c)
        nn.init.uniform_(self.weight, -bound, bound)

    def forward(self, x, edge_index, batch, num_graphs):
        score = (x * self.weight).sum(dim=-1)
        score = torch.tanh(score / self.weight.norm(p=2, dim=-1))
        perm = topk_perm(score.detach(), self.ratio, batch, num_graphs)
        xo = x[perm] * score[perm].view(-1, 1)
        bo = batch[perm]
        n = x.shape[0]
        remap = torch.full((n,), -1, dtype=torch.long)
        remap[perm] = torch.arange(len(perm))
        s, d = remap[edge_index[0]], remap[edge_index[1]]
        m = (s >= 0) & (d >= 0)
        return xo, torch.stack([s[m], d[m]]), bo, perm, score[perm]


def global_max_pool(x, batch, num_graphs):
    out = torch.full((num_graphs, x.shape[1]), -float("inf"), dtype=x.dtype)
    idx = batch[:, None].expand_as(x)
    return out.scatter_reduce(0, idx, x, reduce="amax", include_self=True)


def global_mean_pool(x, batch, num_graphs):
    s = scatter_sum(x, batch, num_graphs)
    cnt = scatter_sum(torch.ones(len(batch), dtype=x.dtype), batch, num_graphs).clamp(min=1)
    return s / cnt[:, None]


def _unpack(data):
    x = data.x.float()
    ei = data.edge_index
    batch = getattr(data, "batch", None)
    if batch is None:
        batch = torch.zeros(x.shape[0], dtype=torch.long)
    ng = int(batch.max()) + 1 if len(batch) else 0
    ng = getattr(data, "num_graphs", ng) or ng
    return x, ei, batch, ng


class NodeRemovalNet(nn.Module):
    """/root/reference/airfoilgcnn.py:24-145 (conv3/pool3/conv6/pool6 exist but are unused)."""

    def __init__(self, output_dim, conv_width=64, topk=0.5, initial_num_nodes=None):
        super().__init__()
        self.conv_width = conv_width
        self.initial_num_nodes = initial_num_nodes
        w = conv_width
        self.conv1 = SAGEConv(2, w)
        self.pool1 = TopKPooling(w, topk)
        self.conv2 = SAGEConv(w, w)
        self.pool2 = TopKPooling(w, topk)
        self.conv3 = SAGEConv(w, w)
        self.pool3 = TopKPooling(w, topk)
        self.conv4 = GCNConv(w, w)
        self.pool4 = TopKPooling(w, topk)
        self.conv5 = GCNConv(w, w)
        self.pool5 = TopKPooling(w, topk)
        self.conv6 = GCNConv(w, w)
        self.pool6 = TopKPooling(w, topk)
        self.lin1 = nn.Linear(2 * w, 128)
        self.lin2 = nn.Linear(128, 64)
        self.lin3 = nn.Linear(64, output_dim)
        torch.manual_seed(0)
        self.reset()

    def reset(self):
        for c in (self.conv1, self.conv2, self.conv3):
            nn.init.xavier_normal_(c.lin_l.weight, gain=0.9)
            nn.init.normal_(c.lin_l.bias)
            nn.init.xavier_normal_(c.lin_r.weight, gain=0.9)
        for c in (self.conv4, self.conv5, self.conv6):
            nn.init.xavier_normal_(c.lin.weight, gain=0.9)
        for l in (self.lin1, self.lin2, self.lin3):
            nn.init.xavier_normal_(l.weight, gain=0.9)
            nn.init.normal_(l.bias)

    def set_num_nodes(self, n):
        self.initial_num_nodes = n
        self.conv1 = SAGEConv(n, self.conv_width)

    def forward(self, data, embedding=False):
        x, ei, batch, ng = _unpack(data)
        acc = None
        for conv, pool in ((self.conv1, self.pool1), (self.conv2, self.pool2),
                           (self.conv4, self.pool4), (self.conv5, self.pool5)):
            x = F.relu(conv(x, ei))
            x, ei, batch, _, _ = pool(x, ei, batch, ng)
            r = torch.cat([global_max_pool(x, batch, ng), global_mean_pool(x, batch, ng)], dim=1)
            acc = r if acc is None else acc + r
        if embedding:
            return acc
        x = F.relu(self.lin1(acc))
        x = F.relu(self.lin2(x))
        x = self.lin3(x)
        return F.softmax(x, dim=1)


class AirfoilGCNN(nn.Module):
    """/root/reference/airfoilgcnn.py:148-209."""

    def __init__(self, conv_width=64):
        super().__init__()
        w, topk = conv_width, 0.5
        self.conv1 = SAGEConv(2, w)
        self.pool1 = TopKPooling(w, topk)
        self.conv2 = SAGEConv(w, w)
        self.pool2 = TopKPooling(w, topk)
        self.conv3 = SAGEConv(w, w)
        self.pool3 = TopKPooling(w, topk)
        self.conv4 = GCNConv(w, w)
        self.pool4 = TopKPooling(w, topk)
        self.conv5 = GCNConv(w, w)
        self.pool5 = TopKPooling(w, topk)
        self.conv6 = GCNConv(w, w)
        self.pool6 = TopKPooling(w, topk)
        self.lin1 = nn.Linear(2 * w, 128)
        self.lin2 = nn.Linear(128, 64)
        self.lin3 = nn.Linear(64, 1)

    def forward(self, data):
        x, ei, batch, ng = _unpack(data)
        x = x[:, [2, 3]]
        acc = None
        for k in range(1, 7):
            conv, pool = getattr(self, f"conv{k}"), getattr(self, f"pool{k}")
            x = F.relu(conv(x, ei))
            x, ei, batch, _, _ = pool(x, ei, batch, ng)
            r = torch.cat([global_max_pool(x, batch, ng), global_mean_pool(x, batch, ng)], dim=1)
            acc = r if acc is None else acc + r
        x = F.relu(self.lin1(acc))
        x = F.relu(self.lin2(x))
        return self.lin3(x)


def replay_loss(net1, net2, states, actions, next_states, rewards, gamma, select=True):
    """DataWorker._get_data + compute_gradients (airfoil_dqn.py:240-310).

    ``states``: Batch of B graphs; ``actions`` i64 [B]; ``next_states``: (Batch of the
    non-final next states, bool mask [B]); ``rewards`` f32 [B].  Returns the Huber
    (delta=1, mean) loss; gradient flows through net1 iff ``select`` else through net2.
    """
    nb, mask = next_states
    B = actions.shape[0]
    with torch.set_grad_enabled(select):
        out = net1(states)
    pred = out[torch.arange(B), actions]
    nsv = torch.zeros(B)
    if mask.any():
        with torch.set_grad_enabled(not select):
            q2 = net2(nb).max(1)[0].float()
        nsv = nsv.index_put((mask.nonzero()[:, 0],), q2)
    target = nsv * gamma + rewards
    return F.huber_loss(pred.float(), target.float())
