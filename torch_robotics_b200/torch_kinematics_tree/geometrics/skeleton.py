"""Link skeleton of a kinematic tree for drawing (reference: geometrics/skeleton.py:133-158 builds a networkx graph from the
FK of one configuration and draws its edges).  Only what examples/inverse_kinematics.py uses: get_skeleton_from_model and
Skeleton.draw_skeleton; the posture-sampling / mediapipe parts of the reference class are outside the scope."""
from __future__ import annotations

import numpy as np
import torch


class Skeleton:
    def __init__(self, node_pos, edges, names):
        self.node_pos = {i: np.asarray(p, np.float64) for i, p in enumerate(node_pos)}
        self.node_name = dict(enumerate(names))
        self.node_ids = list(self.node_pos)
        self.edges = [(int(u), int(v)) for u, v in edges]
        self.three_d = len(node_pos[0]) == 3

    def get_edges(self, weighted=False):
        e = np.array(self.edges)
        if not weighted:
            return e
        w = np.array([[np.linalg.norm(self.node_pos[u] - self.node_pos[v])] for u, v in self.edges])
        return np.concatenate([e, w], axis=1)

    def get_pos_tensor(self):
        return torch.tensor(np.stack([self.node_pos[i] for i in self.node_ids]))

    def draw_skeleton(self, pos=None, shift=None, with_labels=False, color="b", ax=None):
        pos = self.node_pos if pos is None else pos
        shift = np.zeros(3 if self.three_d else 2) if shift is None else np.asarray(shift)
        if ax is None:
            import matplotlib.pyplot as plt
            ax = plt.figure().add_subplot(projection="3d" if self.three_d else None)
        for u, v in self.edges:
            a, b = np.asarray(pos[u]) + shift, np.asarray(pos[v]) + shift
            ax.plot(*[[a[k], b[k]] for k in range(len(a))], color=color)
        pts = np.stack([np.asarray(pos[i]) + shift for i in self.node_ids])
        ax.scatter(*[pts[:, k] for k in range(pts.shape[1])], color=color)
        if with_labels:
            for i in self.node_ids:
                ax.text(*(np.asarray(pos[i]) + shift), self.node_name[i])
        return ax


def get_skeleton_from_model(model, q_calib, link_list=None):
    """Skeleton of `model` at configuration q_calib [n_dofs]: one node per link of link_list (default: all), one edge per
    parent -> child pair whose two ends are both listed."""
    names = list(link_list) if link_list is not None else model.get_link_names()
    q = q_calib.detach().reshape(1, -1)
    H = model.compute_forward_kinematics_all_links(q, link_list=names)[0]
    pos = H[:, :3, 3].detach().cpu().numpy()
    idx = {n: i for i, n in enumerate(names)}
    desc = model._desc
    edges = []
    for n in names:
        b = model._name_to_idx_map[n]
        p = int(desc.parent[b])
        while p >= 0 and desc.link_names[p] not in idx:      # nearest listed ancestor
            p = int(desc.parent[p])
        if p >= 0:
            edges.append((idx[desc.link_names[p]], idx[n]))
    return Skeleton(pos, edges, names)
