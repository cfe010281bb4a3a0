"""Minimal stand-ins for ``torch_geometric.data.Data`` / ``Batch`` / ``DataLoader``.

The reference builds its state as a PyG ``Data(x, edge_index, edge_attr)``
(/root/reference/Env2DAirfoil.py:290) and batches replay transitions with
``torch_geometric.loader.DataLoader`` (/root/reference/airfoil_dqn.py:256,268).
torch_geometric is not available offline, so these duck-typed classes carry the
same attributes (``x``, ``edge_index``, ``edge_attr``, ``batch``, ``ptr``,
``num_graphs``, ``.to(device)``) with PyG's collation rule: node features are
concatenated, ``edge_index`` is offset by the cumulative node count, ``batch``
maps each node to its graph.
"""
from __future__ import annotations

import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, **kw):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.batch = None
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return 0 if self.x is None else int(self.x.shape[0])

    @property
    def num_edges(self):
        return 0 if self.edge_index is None else int(self.edge_index.shape[1])

    def to(self, device, non_blocking=False):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v
        return out

    def pin_memory(self):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.pin_memory() if torch.is_tensor(v) and not v.is_cuda else v
        return out

    def __repr__(self):
        xs = None if self.x is None else list(self.x.shape)
        es = None if self.edge_index is None else list(self.edge_index.shape)
        return f"{self.__class__.__name__}(x={xs}, edge_index={es})"


class Batch(Data):
    """Collated graphs.  ``ptr`` (i64 [B+1], host list mirrored in ``ptr_list``) holds node
    offsets and ``eptr`` edge offsets, so kernels can address one graph per CTA."""

    @classmethod
    def from_data_list(cls, data_list):
        xs, eis, bs = [], [], []
        ptr, eptr = [0], [0]
        for g, d in enumerate(data_list):
            n = d.x.shape[0]
            xs.append(d.x)
            ei = d.edge_index
            if ei.numel() == 0:
                ei = ei.reshape(2, 0).to(torch.long)
            eis.append(ei + ptr[-1])
            bs.append(torch.full((n,), g, dtype=torch.long, device=d.x.device))
            ptr.append(ptr[-1] + n)
            eptr.append(eptr[-1] + ei.shape[1])
        out = cls(x=torch.cat(xs, 0), edge_index=torch.cat(eis, 1))
        out.batch = torch.cat(bs, 0)
        out.ptr = torch.tensor(ptr, dtype=torch.long)
        out.eptr = torch.tensor(eptr, dtype=torch.long)
        out.num_graphs = len(data_list)
        return out

    def _host_meta(self):
        """(ptr i32, eptr i32, B, max_n, max_e) from the HOST offset vectors -- what the kernels need to address one
        graph per CTA; computed once per collated batch so that moving it to a device never synchronises."""
        m = self.__dict__.get("_meta")
        if m is None and getattr(self, "ptr", None) is not None and not self.ptr.is_cuda:
            ptr, eptr = self.ptr, self.eptr
            m = (ptr.to(torch.int32), eptr.to(torch.int32), int(ptr.numel()) - 1,
                 int((ptr[1:] - ptr[:-1]).max()) if ptr.numel() > 1 else 0,
                 int((eptr[1:] - eptr[:-1]).max()) if eptr.numel() > 1 else 0)
            self.__dict__["_meta"] = m
        return m

    def pin_memory(self):
        m = self._host_meta()
        out = super().pin_memory()
        if m is not None:
            out.__dict__["_meta"] = (m[0].pin_memory(), m[1].pin_memory()) + m[2:]
        return out

    def to(self, device, non_blocking=False):
        m = self._host_meta()
        out = super().to(device, non_blocking=non_blocking)
        out.__dict__.pop("_mdq_ptrs", None)
        if m is not None and torch.device(device).type == "cuda":
            out.__dict__["_mdq_ptrs"] = (m[0].to(device, non_blocking=non_blocking), m[1].to(device, non_blocking=non_blocking),
                                         m[2], m[3], m[4])
        return out


class DataLoader:
    """``DataLoader(list_of_Data, batch_size)`` -> iterates ``Batch`` objects (no shuffling by
    default, as the reference uses it: airfoil_dqn.py:256)."""

    def __init__(self, dataset, batch_size=1, shuffle=False):
        self.dataset = list(dataset)
        self.batch_size = batch_size
        self.shuffle = shuffle

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        idx = list(range(len(self.dataset)))
        if self.shuffle:
            idx = torch.randperm(len(idx)).tolist()
        for i in range(0, len(idx), self.batch_size):
            yield Batch.from_data_list([self.dataset[j] for j in idx[i:i + self.batch_size]])
