"""Host-side mirror of hwer/recommendation_base.py: the same classes, method names, argument meaning, return
conventions and error behaviour, with every computation on the path delegated to the CUDA library.

  Node, Edge                       hwer/recommendation_base.py:19-61
  MultiKNN(nodes_to_idx, vectors)  :64-83   per-node-type exact index; query -> [(Node, euclidean dist)] ascending
  RecommendationBase               :86-174  add_nodes / __build_knn__ / fit / predict / get_embeddings /
                                            get_average_embeddings / find_closest_neighbours
plus the entry points BASELINE.json's north_star names (`find_items_for_user`, `find_similar_items`: thin
aliases over find_closest_neighbours) and batched, tensor-returning variants that avoid
materialising Python tuples for large anchor sets.
"""
import abc
import operator
from collections import defaultdict
from typing import Dict, List, Set, Tuple, Union

import numpy as np
import torch

from . import ops
from .logging import getLogger
from .utils import NodeNotFoundException

NodeType = str
NodeExternalId = Union[str, int]
FeatureName = str


class Node:
    """(node_type, node_external_id) value object of hwer/recommendation_base.py:19-36: equal and hashed by that
    pair, external ids compared as strings, repr = the pair.  The pair and its hash are computed once: nodes key
    every host-side dict of this package (a 10 M-node table does tens of millions of lookups per evaluation)."""

    def __init__(self, node_type, node_external_id):
        self.node_type = node_type
        self.node_external_id = str(node_external_id)
        self._pair = (self.node_type, self.node_external_id)
        self._pair_hash = hash(self._pair)
        self._hint = (0, -1)       # (token of the NodeIndex that placed this object, its row): see NodeIndex.rows_of

    def __hash__(self):
        return self._pair_hash

    def __eq__(self, other):
        if isinstance(other, Node):
            return self._pair == other._pair
        return NotImplemented

    def __getstate__(self):
        # string hashes are per process (PYTHONHASHSEED): never carry the cached hash across a pickle
        state = dict(self.__dict__)
        state.pop("_pair_hash", None)
        state.pop("_hint", None)          # index tokens are per process too
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._pair = (self.node_type, self.node_external_id)
        self._pair_hash = hash(self._pair)
        self._hint = (0, -1)

    def __repr__(self):
        return str(self._pair)


class Edge:
    """(src, dst, weight) of hwer/recommendation_base.py:39-61; iterating yields the three fields, so
    `for u, i, r in edges` works as in the reference."""

    def __init__(self, src: Node, dst: Node, weight: float):
        self.src = src
        self.dst = dst
        self.weight = weight
        self.contents = [src, dst, weight]

    def _triple(self):
        return (self.src, self.dst, self.weight)

    def __iter__(self):
        return iter(self.contents)

    def __hash__(self):
        return hash(self._triple())

    def __eq__(self, other):
        if isinstance(other, Edge):
            return self._triple() == other._triple()
        return NotImplemented

    def __repr__(self):
        return "{src: %s, dst: %s, weight: %s}" % (self.src, self.dst, self.weight)


class _InverseView:
    """row -> Node over a NodeIndex: a dict for the nodes that were added as objects, arithmetic for node ranges."""

    def __init__(self, index):
        self._index = index
        self._objects = {v: k for k, v in dict.items(index)}

    def __getitem__(self, row):
        n = self._objects.get(row)
        if n is not None:
            return n
        for node_type, first_row, count in self._index._ranges:
            if first_row <= row < first_row + count:
                return Node(node_type, row - first_row)
        raise KeyError(row)

    def __len__(self):
        return len(self._index)


class NodeIndex(dict):
    """Node -> global row, with the `.inverse` view the reference gets from bidict (row -> Node).

    Catalogues of tens of millions of nodes do not need a Python object per node: `add_range(node_type, count)`
    registers `count` nodes of one type whose external ids are 0 .. count-1 on consecutive rows; lookups of such
    nodes are arithmetic.  Everything else (nodes added as objects) is an ordinary dict entry."""

    _tokens = iter(range(1, 1 << 62))

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._inverse = None
        self._ranges = []            # (node_type, first_row, count)
        self._range_total = 0
        self._token = next(NodeIndex._tokens)

    def place(self, nodes, first_row):
        """Adds `nodes` on consecutive rows and leaves each object a hint of where it lives, so that looking the SAME
        objects up again (the anchors of a batched query are usually the objects the model was fitted with) costs an
        attribute read instead of a hash + equality call per node."""
        tok = self._token
        for i, n in enumerate(nodes, first_row):
            dict.__setitem__(self, n, i)
            n._hint = (tok, i)
        self._inverse = None

    def rows_of(self, nodes):
        """Rows of `nodes` as a list (-1 = unknown node)."""
        tok, get = self._token, self.get
        return [n._hint[1] if n._hint[0] == tok else get(n, -1) for n in nodes]

    def add_range(self, node_type, count):
        first = len(self)
        self._ranges.append((str(node_type), first, int(count)))
        self._range_total += int(count)
        self._inverse = None
        return first

    def _range_row(self, node):
        if not self._ranges or not isinstance(node, Node):
            return None
        ext = node.node_external_id
        if not ext.isdigit() or (len(ext) > 1 and ext[0] == "0"):
            return None
        i = int(ext)
        for node_type, first_row, count in self._ranges:
            if node.node_type == node_type and i < count:
                return first_row + i
        return None

    def __len__(self):
        return dict.__len__(self) + self._range_total

    def __contains__(self, node):
        return dict.__contains__(self, node) or self._range_row(node) is not None

    def __missing__(self, node):                 # dict.__getitem__ falls through to here
        row = self._range_row(node)
        if row is None:
            raise KeyError(node)
        return row

    def get(self, node, default=None):
        row = dict.get(self, node)
        if row is None:
            row = self._range_row(node)
        return default if row is None else row

    def rows_by_type(self):
        """{node_type: ascending numpy array of that type's global rows}."""
        by_type = defaultdict(list)
        for n, i in dict.items(self):
            by_type[n.node_type].append(i)
        out = {t: [np.asarray(r, dtype=np.int64)] for t, r in by_type.items()}
        for node_type, first_row, count in self._ranges:
            out.setdefault(node_type, []).append(np.arange(first_row, first_row + count, dtype=np.int64))
        return {t: np.sort(np.concatenate(parts)) for t, parts in out.items()}

    @property
    def inverse(self):
        if self._inverse is None or len(self._inverse) != len(self):
            self._inverse = _InverseView(self)
        return self._inverse

    def __setitem__(self, key, value):
        self._inverse = None
        super().__setitem__(key, value)

    def update(self, *a, **k):
        self._inverse = None
        super().update(*a, **k)


def _as_device_table(vectors, device):
    if isinstance(vectors, torch.Tensor):
        t = vectors.to(device=device, dtype=torch.float32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(vectors, dtype=np.float32)).to(device)
    return t.contiguous()


class MultiKNN:
    def __init__(self, nodes_to_idx: Dict[Node, int], vectors, leaf_size=128, shadow=None, device=None,
                 max_norm=None, mode="exact"):
        # leaf_size is accepted for signature compatibility; there is no tree.
        if not torch.cuda.is_available():
            raise RuntimeError("hwer_b200 needs a CUDA device (B200): there is no CPU index")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.nodes_to_idx = nodes_to_idx
        self.mode = mode
        assert len(nodes_to_idx) == len(vectors)
        self.table = _as_device_table(vectors, self.device)
        if isinstance(nodes_to_idx, NodeIndex):
            rows_by_type = nodes_to_idx.rows_by_type()
        else:
            rows_by_type: Dict[str, List[int]] = defaultdict(list)
            for n, i in nodes_to_idx.items():
                rows_by_type[n.node_type].append(i)
        self.idxs: Dict[str, np.ndarray] = {}
        self.idxs_dev: Dict[str, torch.Tensor] = {}
        self.offset: Dict[str, int] = {}
        self.knn: Dict[str, ops.TopKIndex] = {}
        self._tables = {}
        for nt, rows in rows_by_type.items():
            rows = np.asarray(rows, dtype=np.int64)
            self.idxs[nt] = rows
            self.idxs_dev[nt] = torch.from_numpy(rows).to(self.device)
            contiguous = len(rows) > 0 and rows[0] + len(rows) - 1 == rows[-1] and np.all(np.diff(rows) == 1)
            if contiguous:
                # a row range of the shared table: no copy (the reference copies per type, :74)
                sub = self.table[int(rows[0]):int(rows[0]) + len(rows)]
                sh = shadow[int(rows[0]):int(rows[0]) + len(rows)] if shadow is not None else None
                self.offset[nt] = int(rows[0])
            else:
                sub = self.table.index_select(0, self.idxs_dev[nt]).contiguous()
                sh = shadow.index_select(0, self.idxs_dev[nt]).contiguous() if shadow is not None else None
                self.offset[nt] = None
            self._tables[nt] = sub
            self.knn[nt] = ops.TopKIndex(sub, sh, max_norm=max_norm)

    def query_batch(self, embeddings, node_type, k=200, mode=None, want_f64=False):
        """[B, d] query embeddings -> (global rows [B, k] int64, dot products [B, k] fp32[, fp64]) on the device,
        ordered (score descending, row ascending).  A type whose rows are a range of the table gets its global
        rows straight from the search (idx_offset); a gathered type goes through hwer_map_rows."""
        q = _as_device_table(embeddings, self.device)
        if q.dim() == 1:
            q = q[None, :]
        off = self.offset[node_type]
        res = self.knn[node_type].topk(q, k, mode or self.mode, idx_offset=off or 0, want_f64=want_f64)
        rows = res[0] if off is not None else ops.map_rows(res[0], self.idxs_dev[node_type])
        return (rows,) + tuple(res[1:])

    def query(self, embedding, node_type, k=200) -> List[Tuple[Node, float]]:
        """hwer/recommendation_base.py:78-83: k nearest rows of `node_type` by Euclidean distance, ascending
        (distance of the fp32 rows to the embedding in float64, like KDTree64)."""
        q = _as_device_table(embedding, self.device).reshape(1, -1)
        rows, _ = self.query_batch(q, node_type, k=k)
        rows, dist = ops.rerank(self.table, rows, "euclid", queries=q)
        inv = self.nodes_to_idx.inverse
        return [(inv[i], dt) for i, dt in zip(rows[0].cpu().tolist(), dist[0].cpu().tolist()) if i >= 0]


class RecommendationBase(metaclass=abc.ABCMeta):
    def __init__(self, node_types: Set[str], n_dims: int = 32, device=None, mode: str = "exact"):
        self.node_types: Set[NodeType] = node_types
        self.nodes_to_idx: NodeIndex = NodeIndex()
        self.knn: MultiKNN = None
        self.vectors = None            # what __build_knn__ was given (numpy in the reference)
        self.device_vectors = None     # the resident fp32 table the kernels read
        self.fit_done = False
        self.n_dims = n_dims
        self.device = device
        self.mode = mode               # "exact" (fp64-rescored, the reference's result) or "bf16"
        self.log = getLogger(type(self).__name__)

    def add_node_range(self, node_type: str, count: int):
        """`count` nodes of `node_type` with external ids 0 .. count-1 on the next `count` rows, without a Python
        object per node (catalogues of tens of millions of items).  Returns the first row."""
        assert node_type in self.node_types
        return self.nodes_to_idx.add_range(node_type, count)

    def add_nodes(self, nodes: List[Node]):
        assert len(set(nodes)) == len(nodes)
        assert not any(n in self.nodes_to_idx for n in nodes)
        assert len(set([n.node_type for n in nodes]) - self.node_types) == 0
        self.nodes_to_idx.place(nodes, len(self.nodes_to_idx))
        return self

    def __build_knn__(self, vectors, shadow=None):
        if not torch.cuda.is_available():
            raise RuntimeError("hwer_b200 needs a CUDA device (B200): there is no CPU index")
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
        table = _as_device_table(vectors, dev)
        v, _, _, _, max_norm = ops.norm_stats(table)
        assert v == 0
        self.knn = MultiKNN(self.nodes_to_idx, table, leaf_size=128, shadow=shadow, device=dev, max_norm=max_norm,
                            mode=self.mode)
        self.vectors = vectors
        self.device_vectors = self.knn.table
        return self

    @abc.abstractmethod
    def fit(self, nodes: List[Node], edges: List[Edge], node_data: Dict[Node, Dict[FeatureName, object]], **kwargs):
        assert not self.fit_done
        edge_node_types = set([node.node_type for e in edges for node in [e.src, e.dst]])
        sparsity = 1 - len(edges) / (len(nodes) * len(nodes))
        self.log.info("Start Fitting Base Recommender with nodes = %s, edges = %s, sparsity = %s",
                      len(nodes), len(edges), sparsity)
        assert edge_node_types == self.node_types
        assert len(set([i for e in edges for i in [e.src, e.dst]]) - set(nodes)) == 0
        assert len(set(nodes)) == len(nodes)
        assert len(set([n.node_type for n in nodes]) - self.node_types) == 0
        self.add_nodes(nodes)
        self.log.info("End Fitting Base Recommender")
        return edges

    # ------------------------------------------------------------------ pair scores
    def _rows_to_device(self, rows) -> torch.Tensor:
        """A Python list of rows -> int64 device tensor through a pinned staging buffer (grow-only, reused): numpy
        fills it in one pass and the copy is asynchronous -- torch.tensor(list, device=...) walks the list
        element by element and copies from pageable memory (0.23 ms for 4096 rows, profiles/r02_q_time_api.txt)."""
        n = len(rows)
        st = getattr(self, "_row_staging", None)
        if st is None or st.shape[0] < n:
            st = torch.empty((max(n, 4096),), dtype=torch.int64).pin_memory()
            self._row_staging = st
            self._row_staging_ev = None
        elif self._row_staging_ev is not None:
            self._row_staging_ev.synchronize()          # the previous upload has left the buffer
        st.numpy()[:n] = np.fromiter(rows, dtype=np.int64, count=n)
        dev = self.device_vectors.device
        out = st[:n].to(dev, non_blocking=True)
        self._row_staging_ev = torch.cuda.Event()
        self._row_staging_ev.record(torch.cuda.current_stream(dev))
        return out

    def _rows_of(self, nodes) -> torch.Tensor:
        return self._rows_to_device(self.nodes_to_idx.rows_of(nodes))

    def predict_rows(self, src_rows: torch.Tensor, dst_rows: torch.Tensor) -> torch.Tensor:
        return ops.pair_score(self.device_vectors, src_rows, dst_rows)

    def predict(self, node_pairs: List[Tuple[Node, Node]]) -> List[float]:
        """Probability-like link score (dot + 1) / 2 for each pair; unknown nodes score ~0.5 (:135-151)."""
        src, dst = zip(*node_pairs)
        return self.predict_rows(self._rows_of(src), self._rows_of(dst)).cpu().numpy()

    def get_embeddings(self, nodes: List[Node]):
        """:146-151: the nodes' rows; a node never trained on gets clip(row 0, 1e-6, 1e-5)."""
        return ops.gather_rows(self.device_vectors, self._rows_of(nodes)).cpu().numpy()

    def get_average_embeddings(self, entities: List[Node]):
        """:153-155: unit(mean(rows of the entities))."""
        return self._average_embedding(self._rows_of(entities)).cpu().numpy()

    def _average_embedding(self, rows: torch.Tensor) -> torch.Tensor:
        ptr = torch.tensor([0, rows.shape[0]], dtype=torch.int64, device=rows.device)
        return ops.average_embeddings(self.device_vectors, ptr, rows)[0]

    def _csr_rows(self, lists):
        """Per-anchor node lists -> (ptr [B+1], rows) device CSR for ops.compose_queries."""
        ptr, rows = [0], []
        for nodes in lists:
            rows.extend(self.nodes_to_idx[n] if n in self.nodes_to_idx else -1 for n in (nodes or []))
            ptr.append(len(rows))
        dev = self.device_vectors.device
        return (torch.tensor(ptr, dtype=torch.int64, device=dev), torch.tensor(rows, dtype=torch.int64, device=dev))

    def _query_embeddings(self, anchors, positive=None, negative=None) -> torch.Tensor:
        """[B, d] query vectors of :164-170 for a batch of anchors; positive / negative: None or one node list per
        anchor (empty or None entries = that part is absent for the anchor)."""
        pos = self._csr_rows(positive) if positive is not None and any(positive) else None
        neg = self._csr_rows(negative) if negative is not None and any(negative) else None
        return ops.compose_queries(self.device_vectors, self._rows_of(anchors), pos, neg)

    def _query_embedding(self, anchor, positive=None, negative=None) -> torch.Tensor:
        return self._query_embeddings([anchor], [positive] if positive else None, [negative] if negative else None)[0]

    # ------------------------------------------------------------------ retrieval
    def _check_query(self, node_type, anchor):
        assert self.fit_done
        assert node_type in self.node_types and node_type in self.knn.knn
        if anchor not in self.nodes_to_idx:
            raise NodeNotFoundException("Node = %s, was not provided in training" % anchor)

    def find_closest_neighbours(self, node_type: str, anchor: Node, positive: List[Node] = None,
                                negative: List[Node] = None, k=200) -> List[Tuple[Node, float]]:
        """:157-174 (and the GcnNCF override, hwer/gcn_ncf.py:363-387, through `_batch_scores`): one anchor is a
        batch of one -- search, score convention and ordering all run in the same kernels as the batched call."""
        self._check_query(node_type, anchor)
        rows, scores = self.find_closest_neighbours_batch(node_type, [anchor], k=k,
                                                          positive=[positive] if positive else None,
                                                          negative=[negative] if negative else None)
        return self.rows_to_nodes(rows, scores)[0]

    def find_items_for_user(self, user: Node, k=200, positive: List[Node] = None, negative: List[Node] = None,
                            node_type: str = "item") -> List[Tuple[Node, float]]:
        """north_star name for find_closest_neighbours(node_type='item', anchor=<user>)."""
        return self.find_closest_neighbours(node_type, user, positive, negative, k)

    def find_similar_items(self, item: Node, k=200, positive: List[Node] = None, negative: List[Node] = None
                           ) -> List[Tuple[Node, float]]:
        """north_star name for find_closest_neighbours(node_type=<item's type>, anchor=<item>); like the
        reference, the anchor itself is returned first."""
        return self.find_closest_neighbours(item.node_type, item, positive, negative, k)

    def _batch_scores(self, anchor_rows, queries, rows):
        """Final score convention of the base class (:172-174): (anchor . node + 1) / 2 of the table's fp32 rows --
        NOT the composed query's score -- each anchor's list sorted descending (stable)."""
        return ops.rerank(self.device_vectors, rows, "pair", anchor_rows=anchor_rows)

    def find_closest_neighbours_batch(self, node_type: str, anchors, k=200,
                                      positive: List[List[Node]] = None, negative: List[List[Node]] = None
                                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """One search for all anchors (the loop of validation.model_get_topk_knn, hwer/validation.py:30-35).
        Returns device tensors (global rows [B, k] int64, scores [B, k] float64) in the same order and score
        convention as find_closest_neighbours(node_type, anchor, positive[i], negative[i], k) called per anchor."""
        assert self.fit_done
        assert node_type in self.node_types and node_type in self.knn.knn
        if isinstance(anchors, torch.Tensor):
            # anchors already resolved to global rows (a serving tier that keeps its own id map): int64, any device
            anchor_rows = anchors.to(device=self.device_vectors.device, dtype=torch.int64, non_blocking=True)
        else:
            rows = self.nodes_to_idx.rows_of(anchors)
            if rows and min(rows) < 0:
                raise NodeNotFoundException("Node = %s, was not provided in training" % anchors[rows.index(min(rows))])
            anchor_rows = self._rows_to_device(rows)
        pos = self._csr_rows(positive) if positive is not None and any(positive) else None
        neg = self._csr_rows(negative) if negative is not None and any(negative) else None
        queries = ops.compose_queries(self.device_vectors, anchor_rows, pos, neg)
        rows, _ = self.knn.query_batch(queries, node_type, k=k)
        return self._batch_scores(anchor_rows, queries, rows)

    def find_closest_neighbours_batch_to_host(self, node_type: str, anchors, k=200, out=None, chunk: int = 32768
                                              ) -> Tuple[torch.Tensor, torch.Tensor]:
        """find_closest_neighbours_batch for very large anchor sets (all users of a catalogue), delivering into PINNED
        host memory: the anchors are answered in chunks and chunk i's rows / scores travel to the host on a second
        stream while chunk i + 1 is searched, so the result copy (16 bytes per neighbour: 222 MB for 138,493 users x
        100) overlaps the search instead of following it.  `out`: optional (rows [B, k] int64, scores [B, k] float64)
        pinned tensors to fill; returns them after a synchronisation.  Same rows, scores and order as the plain call."""
        n = anchors.shape[0] if isinstance(anchors, torch.Tensor) else len(anchors)
        if out is None:
            out = (torch.empty((n, k), dtype=torch.int64).pin_memory(), torch.empty((n, k), dtype=torch.float64).pin_memory())
        rows_h, sc_h = out
        dev = self.device_vectors.device
        main = torch.cuda.current_stream(dev)
        copier = getattr(self, "_copy_stream", None)
        if copier is None:
            copier = self._copy_stream = torch.cuda.Stream(device=dev)
        keep = []                                       # chunk results stay referenced until their copies are done
        for b in range(0, n, chunk):
            e = min(n, b + chunk)
            r, s = self.find_closest_neighbours_batch(node_type, anchors[b:e], k=k)
            ready = torch.cuda.Event()
            ready.record(main)
            copier.wait_event(ready)
            with torch.cuda.stream(copier):
                rows_h[b:e].copy_(r, non_blocking=True)
                sc_h[b:e].copy_(s, non_blocking=True)
            keep.append((r, s))
        copier.synchronize()
        main.synchronize()
        return rows_h, sc_h

    def rows_to_nodes(self, rows: torch.Tensor, scores: torch.Tensor) -> List[List[Tuple[Node, float]]]:
        inv = self.nodes_to_idx.inverse
        return [[(inv[i], s) for i, s in zip(r, sc) if i >= 0] for r, sc in zip(rows.cpu().tolist(), scores.cpu().tolist())]
