"""Serving-side mirrors of the reference's two recommenders.

The reference classes train their tables inside fit() (feature encoders + PCA for ContentRecommendation,
DGL GraphSAGE + optional NCF for GcnNCF); training is out of scope here.  These
classes take the trained [N, d] tables as fit() keyword arguments and reproduce everything from the blend slot
onwards:

  ContentRecommendation.fit tail      hwer/content_recommender.py:94-97   (__build_knn__, fit_done, return table)
  GcnNCF.prepare_for_knn              hwer/gcn_ncf.py:447-456             (+ the north_star alpha blend)
  GcnNCF.fit tail                     hwer/gcn_ncf.py:439-445
  GcnNCF.predict (cosine branch)      hwer/gcn_ncf.py:330-334
  GcnNCF.find_closest_neighbours      hwer/gcn_ncf.py:363-383             score = (2 - euclidean distance) / 2
  GcnNCF.predict (NCF branch)         hwer/gcn_ncf.py:336-361             sigmoid(MLP([h_src || h_dst])), hwer/ncf.py:7-27
  GcnNCF.find_closest_neighbours      hwer/gcn_ncf.py:384-386             NCF re-rank of the k retrieved nodes
"""
import operator
from typing import Dict, List, Set, Tuple

import numpy as np
import torch

from . import ops
from .recommendation_base import Edge, FeatureName, Node, RecommendationBase, _as_device_table


class ContentRecommendation(RecommendationBase):
    def __init__(self, embedding_mapper=None, node_types: Set[str] = None, n_dims: int = 32, device=None,
                 mode: str = "exact"):
        super().__init__(node_types, n_dims, device=device, mode=mode)
        self.embedding_mapper = embedding_mapper

    def fit(self, nodes: List[Node], edges: List[Edge], node_data: Dict[Node, Dict[FeatureName, object]] = None,
            vectors=None, hyperparameters=None, **kwargs):
        """`vectors`: the finished [N, n_dims] unit-norm content table (what __build_content_embeddings__
        produces in the reference), row i belonging to nodes[i].  It may also arrive as hyperparameters["vectors"]
        (the keyword validation.test_algorithm passes, hwer/validation.py:197,202)."""
        if vectors is None and hyperparameters:
            vectors = hyperparameters.get("vectors")
        if vectors is None:
            raise ValueError("hwer_b200 serves trained tables: pass vectors=<[N, d] unit-norm array>")
        super().fit(nodes, edges, node_data, **kwargs)
        self.__build_knn__(vectors)
        self.fit_done = True
        return self.vectors


class GcnNCF(RecommendationBase):
    def __init__(self, embedding_mapper=None, node_types: Set[str] = None, n_dims: int = 32, alpha: float = 0.0,
                 device=None, mode: str = "exact"):
        super().__init__(node_types, n_dims, device=device, mode=mode)
        assert n_dims % 2 == 0
        self.embedding_mapper = embedding_mapper
        self.alpha = alpha            # weight of the content table in the blend; 0.0 = the reference's behaviour
        self.ncf_enabled = False      # set by set_ncf(): the reference sets it when ncf_epochs > 0 (gcn_ncf.py:437)
        self.prediction_artifacts = dict()
        self.shadow = None

    def set_ncf(self, h, params, depth):
        """Installs a trained NCF re-ranker (the reference's prediction_artifacts {"model", "h"}, gcn_ncf.py:320-322).
        h: [N + 1, n_dims] NCF input vectors, row 0 = the padding node; params: the flat fp32 parameter vector
        [W1, b1, ..., W_depth, b_depth, w_out, b_out] in torch Linear layout (e.g. concatenated from
        `model.W` of hwer/ncf.py); depth: ncf_layers."""
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
        h = _as_device_table(h, dev)
        if isinstance(params, np.ndarray):
            params = torch.from_numpy(np.ascontiguousarray(params, dtype=np.float32))
        params = params.to(dev, torch.float32).contiguous()
        if params.shape[0] != ops.ncf_param_count(h.shape[1], depth):
            raise ValueError("NCF parameter vector does not match width %d, depth %d" % (h.shape[1], depth))
        self.prediction_artifacts = {"h": h, "params": params, "depth": int(depth)}
        self.ncf_enabled = True
        return self

    def get_gcn_vectors(self, gcn_params: Dict[str, object], content_vectors, edges: List[Edge] = None,
                        neighbours=None, seed: int = 0, previous=None, ema: float = 0.1) -> torch.Tensor:
        """The inference pass of the reference's get_gcn_vectors (hwer/gcn_ncf.py:260-279): the trained
        GraphConvModule (hwer/gcn.py:146-193) applied to every node, producing the collaborative vectors that
        prepare_for_knn turns into the serving table.  `gcn_params`: the module's tensors by name -- node_emb
        [(n + 1), F], proj_w [F, C], proj_b, ln_g, ln_b [F], fc0_w [4F, F (layers + 1)], fc0_b, fc1_w [F, 4F], fc1_b
        (torch state_dict entries `node_emb.weight`, `proj.0.*`, `proj.2.*`, `convs.<L-1>.fc.0.*`, `convs.<L-1>.fc.3.*`).
        The neighbour sample of every block is either given (`neighbours`: one (ptr, idx) CSR pair per GCN layer) or
        drawn here from `edges` like the reference's sampler does (two random in-neighbours + a self loop per node and
        layer, seeded).  `previous`: the module's EMA state, updated in place."""
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)

        def t(x, dtype=torch.float32):
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(np.ascontiguousarray(x))
            return x.to(dev, dtype).contiguous()

        g = {k_: t(v_) for k_, v_ in gcn_params.items()}
        content = t(content_vectors)
        layers = g["fc0_w"].shape[1] // g["node_emb"].shape[1] - 1
        if neighbours is None:
            assert edges is not None, "either `neighbours` or `edges` is needed"
            src = self.nodes_to_idx.rows_of([e.src for e in edges])
            dst = self.nodes_to_idx.rows_of([e.dst for e in edges])
            neighbours = ops.sample_neighbours(content.shape[0], src, dst, fanout=2, seed=seed, blocks=layers)
        nbr = [(t(p_, torch.int64), t(i_, torch.int64)) for p_, i_ in neighbours]
        prev = None
        if previous is not None:
            prev = previous if isinstance(previous, torch.Tensor) and previous.is_cuda else t(previous)
        return ops.gcn_infer(g["node_emb"], content, g["proj_w"], g["proj_b"], g["ln_g"], g["ln_b"], nbr, g["fc0_w"],
                             g["fc0_b"], g["fc1_w"], g["fc1_b"], previous=prev, ema=ema)

    def predict_rows(self, src_rows: torch.Tensor, dst_rows: torch.Tensor) -> torch.Tensor:
        if not self.ncf_enabled:
            return super().predict_rows(src_rows, dst_rows)
        pa = self.prediction_artifacts
        # gcn_ncf.py:341-342: node row + 1, unknown node -> 0 (the padding row)
        return ops.ncf_score(pa["h"], pa["params"], (src_rows + 1).clamp(min=0), (dst_rows + 1).clamp(min=0), pa["depth"])

    def predict(self, node_pairs):
        res = super().predict(node_pairs)
        return list(res) if self.ncf_enabled else res      # the reference returns a list here (gcn_ncf.py:360-361)

    def _pca_reduce(self, vectors, dev):
        """PCA(n_components=n_dims).fit_transform of hwer/gcn_ncf.py:449-452 on the device: centre, covariance
        (one cuBLAS GEMM), symmetric eigendecomposition (cuSOLVER), projection on the n_dims leading axes, with
        sklearn's sign rule (each axis' largest-magnitude loading is positive).  An offline step of fit(), run once
        per table -- library calls, not hot-path kernels.  The reference's call leaves the solver to sklearn's 'auto'
        (randomised and unseeded for large tables), so its own output is only defined up to that solver's error; this
        is the exact ('full') decomposition in float64, rounded to fp32 like the reference's result."""
        x = vectors if isinstance(vectors, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(vectors))
        x = x.to(dev, torch.float64)
        x = x - x.mean(dim=0, keepdim=True)
        evals, evecs = torch.linalg.eigh(x.t() @ x)                   # ascending eigenvalues
        axes = evecs[:, -self.n_dims:].flip(1)                        # [D, n_dims], leading axis first
        lead = axes.abs().argmax(dim=0)
        axes = axes * torch.sign(axes[lead, torch.arange(axes.shape[1], device=dev)])[None, :]
        return (x @ axes).float().contiguous()

    def prepare_for_knn(self, content_vectors, collaborative_vectors, alpha=None):
        """unit(alpha * unit(content) + (1 - alpha) * unit(collaborative)) on the device; numpy in -> numpy out."""
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
        as_tensor = isinstance(collaborative_vectors, torch.Tensor)      # tensor in -> tensor out, numpy in -> numpy out
        if collaborative_vectors.shape[1] > self.n_dims:
            collaborative_vectors = self._pca_reduce(collaborative_vectors, dev)        # gcn_ncf.py:449-452
        elif collaborative_vectors.shape[1] < self.n_dims:
            raise ValueError()
        alpha = self.alpha if alpha is None else alpha
        g = _as_device_table(collaborative_vectors, dev)
        use_content = content_vectors is not None and not (isinstance(alpha, float) and alpha == 0.0)
        c = None
        if use_content:
            if content_vectors.shape != collaborative_vectors.shape:
                raise ValueError("content and collaborative tables must have the same shape to be blended")
            c = _as_device_table(content_vectors, dev)
        if isinstance(alpha, np.ndarray):
            alpha = torch.from_numpy(alpha.astype(np.float32)).to(dev)
        table, self.shadow = ops.blend_normalize(c, g, alpha if use_content else 0.0)
        self._device_table = table
        if as_tensor:
            return table
        return table.cpu().numpy()

    def fit(self, nodes: List[Node], edges: List[Edge], node_data: Dict[Node, Dict[FeatureName, object]] = None,
            content_vectors=None, collaborative_vectors=None, alpha=None, hyperparameters=None, **kwargs):
        if hyperparameters:      # the keyword validation.test_algorithm passes (hwer/validation.py:197,202)
            content_vectors = hyperparameters.get("content_vectors") if content_vectors is None else content_vectors
            if collaborative_vectors is None:
                collaborative_vectors = hyperparameters.get("collaborative_vectors")
            alpha = hyperparameters.get("alpha") if alpha is None else alpha
        if collaborative_vectors is None:
            raise ValueError("hwer_b200 serves trained tables: pass collaborative_vectors=<[N, d] array>")
        super().fit(nodes, edges, node_data, **kwargs)
        knn_vectors = self.prepare_for_knn(content_vectors, collaborative_vectors, alpha)
        # hand the device-resident table and its bf16 shadow straight to the index (no second upload)
        RecommendationBase.__build_knn__(self, self._device_table, shadow=self.shadow)
        self.vectors = knn_vectors
        self.fit_done = True
        return self.vectors

    def _batch_scores(self, anchor_rows, queries, rows):
        if self.ncf_enabled:
            # gcn_ncf.py:384-386: one NCF pass over all B * k (anchor, candidate) pairs, then the per-anchor
            # descending sort
            B, k = rows.shape
            given = self.predict_rows(anchor_rows[:, None].expand(B, k).reshape(-1).contiguous(),
                                      rows.reshape(-1)).reshape(B, k)
            return ops.rerank(self.device_vectors, rows, "given", given=given)
        # gcn_ncf.py:378-383: (2 - dist) / 2 with dist = KDTree64's distance of the row to the COMPOSED embedding
        # (anchor +- positive / negative means, :369-376), not to the anchor's own row
        return ops.rerank(self.device_vectors, rows, "dist", queries=queries)
