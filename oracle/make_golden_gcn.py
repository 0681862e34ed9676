"""TEST INFRASTRUCTURE ONLY.  Golden vectors for GCN inference (`get_gcn_vectors`,
hwer/gcn_ncf.py:260-279), produced by running the UNMODIFIED reference module -- GraphConvModule.forward,
hwer/gcn.py:162-193, with its own build_content_layer / GraphConv sub-modules and its own initialisation -- in this
container through oracle/ref_shim.py.

DGL 0.4's NodeFlow / NeighborSampler cannot be installed, so the module is driven by a STAND-IN NodeFlow (below) that
implements exactly the calls forward() makes (copy_from_parent, num_layers, layer_parent_nid, layer_size, layers[i].data,
block_compute, map_from_parent_nid) over EXPLICIT per-layer neighbour lists: what the sampler would have drawn is an
input.  block_compute follows DGL's semantics for the reference's message / reduce pair (copy_src + sum, gcn.py:159-160):
h_agg[v] = sum of the source rows in list order, w[v] = their count, then the layer's apply function.
What stays unpinned is therefore only DGL's random sampler itself; the arithmetic of every layer is the reference's.

    python oracle/make_golden_gcn.py        ->  tests/golden/reference_gcn.npz
"""
import importlib
import os
import types

import numpy as np
import torch

import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Layer:
    def __init__(self):
        self.data = {}


class StandInNodeFlow:
    """Full-graph NodeFlow: every layer holds all n nodes in id order; block i joins layer i -> i + 1 along nbr[i]
    (ptr [n + 1], idx) = the sampled in-neighbours of every node for that block, self loop included by the caller."""

    def __init__(self, n, content, nbr):
        self.n = n
        self.content = content
        self.nbr = nbr
        self.num_layers = len(nbr) + 1
        self.layers = [_Layer() for _ in range(self.num_layers)]

    def copy_from_parent(self, edge_embed_names=None):
        for l in self.layers:
            l.data["content"] = self.content

    def layer_parent_nid(self, i):
        return torch.arange(self.n)

    def layer_size(self, i):
        return self.n

    def map_from_parent_nid(self, layer, ids, remap):
        return ids

    def block_compute(self, i, msg, red, apply_fn):
        ptr, idx = self.nbr[i]
        src = self.layers[i].data
        dst = self.layers[i + 1].data
        F = src["h"].shape[1]
        h_agg = torch.zeros((self.n, F), dtype=src["h"].dtype)
        w = torch.zeros((self.n,), dtype=src["one"].dtype)
        for v in range(self.n):
            for u in idx[ptr[v]:ptr[v + 1]]:
                h_agg[v] += src["h"][u]
                w[v] += src["one"][u]
        nodes = types.SimpleNamespace(data={"h_agg": h_agg, "h": dst["h"], "w": w})
        dst.update(apply_fn(nodes))


def sample_lists(rs, n, edges, fanout):
    """`fanout` random in-neighbours per node (with replacement when there are fewer) + a self loop: the shape of
    NeighborSampler(g, batch, 2, layers, add_self_loop=True) output (gcn_ncf.py:262-272)."""
    inn = [[] for _ in range(n)]
    for a, b in edges:
        inn[b].append(a)
        inn[a].append(b)
    ptr, idx = [0], []
    for v in range(n):
        pick = list(rs.choice(inn[v], size=min(fanout, len(inn[v])), replace=False)) if inn[v] else []
        pick.append(v)
        idx.extend(int(x) for x in pick)
        ptr.append(len(idx))
    return np.asarray(ptr, dtype=np.int64), np.asarray(idx, dtype=np.int64)


def main():
    ref_shim.load_reference()
    gcn = importlib.import_module("hwer.gcn")
    out = {}
    cases = []
    for ci, (n, C, F, L, seed) in enumerate([(300, 48, 32, 2, 21), (200, 20, 64, 1, 22), (150, 36, 16, 3, 23)]):
        torch.manual_seed(seed)
        rs = np.random.RandomState(seed)
        content = torch.from_numpy(rs.standard_normal((n, C)).astype(np.float32))
        edges = [(int(rs.randint(n)), int(rs.randint(n))) for _ in range(4 * n)]
        fake_g = types.SimpleNamespace(number_of_nodes=lambda n=n: n)
        model = gcn.GraphConvModule(C, F, L, fake_g, 0.1)
        # give the EMA state and LayerNorm non-trivial values (both are parameters the trained model carries)
        with torch.no_grad():
            model.previous.copy_(torch.from_numpy(rs.standard_normal((n + 1, F)).astype(np.float32) * 0.1))
            model.proj[2].weight.copy_(torch.from_numpy(1.0 + 0.1 * rs.standard_normal(F).astype(np.float32)))
            model.proj[2].bias.copy_(torch.from_numpy(0.1 * rs.standard_normal(F).astype(np.float32)))
        model.eval()
        nbr = [sample_lists(rs, n, edges, 2) for _ in range(L)]
        prev_before = model.previous.detach().clone().numpy()
        nf = StandInNodeFlow(n, content, [(p.tolist(), i.tolist()) for p, i in nbr])
        with torch.no_grad():
            h = model(nf).detach().numpy().copy()
        pre = "c%d_" % ci
        out[pre + "shape"] = np.array([n, C, F, L])
        out[pre + "content"] = content.numpy()
        out[pre + "node_emb"] = model.node_emb.weight.detach().numpy()
        out[pre + "proj_w"] = model.proj[0].weight.detach().numpy()
        out[pre + "proj_b"] = model.proj[0].bias.detach().numpy()
        out[pre + "ln_g"] = model.proj[2].weight.detach().numpy()
        out[pre + "ln_b"] = model.proj[2].bias.detach().numpy()
        fc = model.convs[L - 1].fc
        out[pre + "fc0_w"] = fc[0].weight.detach().numpy()
        out[pre + "fc0_b"] = fc[0].bias.detach().numpy()
        out[pre + "fc1_w"] = fc[3].weight.detach().numpy()
        out[pre + "fc1_b"] = fc[3].bias.detach().numpy()
        for l, (p, i) in enumerate(nbr):
            out[pre + "nbr_ptr%d" % l] = p
            out[pre + "nbr_idx%d" % l] = i
        out[pre + "previous_before"] = prev_before
        out[pre + "previous_after"] = model.previous.detach().numpy().copy()
        out[pre + "h"] = h
        cases.append(ci)
    out["cases"] = np.asarray(cases)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "reference_gcn.npz"), **out)
    print("wrote", os.path.join(OUT, "reference_gcn.npz"), {k: v.shape for k, v in out.items() if k.startswith("c0_")})


if __name__ == "__main__":
    main()
