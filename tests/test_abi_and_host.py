"""CPU: the C-ABI library loads and exports every symbol include/hwer_b200.h declares; the host-side mirror keeps
the reference's semantics; nothing silently falls back to the CPU."""
import os
import sys
import re

import numpy as np
import pytest
import torch

import hwer_b200
from hwer_b200 import _native
from hwer_b200.recommendation_base import Edge, Node, NodeIndex, RecommendationBase

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "hwer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hwer_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    declared = _header_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), "libhwer_b200.so does not export %s" % name
    assert sorted(_native.SIGNATURES) == declared, "ctypes table and header disagree"
    assert lib.hwer_version() == 100
    assert lib.hwer_shadow_width(1) == 64 and lib.hwer_shadow_width(128) == 128 and lib.hwer_shadow_width(129) == 192


def test_library_is_sm100a_tcgen05_code():
    """The shipped kernels are Blackwell-native: tcgen05 MMA, TMEM loads and TMA appear in the SASS."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _native.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic


def test_argument_validation_without_gpu():
    lib = _native.lib()
    assert lib.hwer_topk(None, None, 1, 1, 0, 0, 0, None, None, None, None) == _native.HWER_E_INVALID
    assert b"hwer_topk" in lib.hwer_last_error()
    assert lib.hwer_pair_score(None, 1, 1, None, None, 1, None, None) == _native.HWER_E_INVALID
    assert lib.hwer_index_destroy(None) == 0


def test_no_cpu_fallback():
    t = torch.zeros(4, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        hwer_b200.ops.norm_stats(t)
    with pytest.raises(RuntimeError, match="no CPU path"):
        hwer_b200.ops.blend_normalize(None, t)
    if not torch.cuda.is_available():
        class R(RecommendationBase):
            def fit(self, *a, **k):
                pass
        r = R({"user"}, 8)
        r.add_nodes([Node("user", i) for i in range(4)])
        with pytest.raises(RuntimeError, match="CUDA"):
            r.__build_knn__(np.eye(4, 8, dtype=np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hybrid-weighted-embedding-recommender_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "hwer_oracle" not in text and "ref_shim" not in text, f
                assert "/root/reference" not in text, f


def test_node_edge_semantics():
    # hwer/recommendation_base.py:19-61
    assert Node("user", 1) == Node("user", "1") and hash(Node("user", 1)) == hash(Node("user", "1"))
    assert Node("user", 1) != Node("item", 1)
    assert repr(Node("item", 7)) == "('item', '7')"
    e = Edge(Node("user", 1), Node("item", 2), 4.0)
    u, i, w = e
    assert (u, i, w) == (Node("user", 1), Node("item", 2), 4.0)
    assert e == Edge(Node("user", "1"), Node("item", "2"), 4.0) and len({e, Edge(u, i, 4.0)}) == 1


def test_add_nodes_contract():
    class R(RecommendationBase):
        def fit(self, *a, **k):
            pass
    r = R({"user", "item"}, 8)
    users = [Node("user", i) for i in range(3)]
    r.add_nodes(users)
    r.add_nodes([Node("item", 0)])
    assert r.nodes_to_idx[Node("item", 0)] == 3 and r.nodes_to_idx.inverse[1] == Node("user", 1)
    with pytest.raises(AssertionError):
        r.add_nodes([Node("user", 0)])                # already present   (:98)
    with pytest.raises(AssertionError):
        r.add_nodes([Node("item", 5), Node("item", 5)])   # duplicates    (:97)
    with pytest.raises(AssertionError):
        r.add_nodes([Node("genre", 1)])               # unknown type      (:99)
    with pytest.raises(AssertionError):
        r.find_closest_neighbours("item", users[0])   # not fitted        (:159)


def test_node_index_inverse_tracks_updates():
    ix = NodeIndex()
    ix.update({Node("a", 1): 0})
    assert ix.inverse[0] == Node("a", 1)
    ix[Node("a", 2)] = 1
    assert ix.inverse[1] == Node("a", 2)


def test_partition_covers_rows_exactly():
    from hwer_b200.sharded import partition
    for n in (1, 7, 128, 10_000_000, 500_000_000):
        for w in (1, 2, 4, 8):
            spans = [partition(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_sampled_negatives_equal_sampling_from_the_sorted_pool():
    """validation.ncf_eval draws its 100 negatives like random.sample(sorted(items - seen, key=repr), 100) (what the
    reference's random.sample(set) amounts to under oracle/ref_shim.py), but from a numpy index vector: same
    consumption of the `random` stream, same picks."""
    import random
    from collections import defaultdict
    import hwer_b200 as hw
    items = [hw.Node("item", i) for i in range(300)] + [hw.Node("item", "x%d" % i) for i in range(20)]
    users = [hw.Node("user", i) for i in range(12)]
    rnd = random.Random(3)
    interactions = defaultdict(set)
    val = []
    for u in users:
        interactions[u] = set(rnd.sample(items, rnd.randint(1, 150))) | {hw.Node("genre", "g")}
        for it in rnd.sample(sorted(interactions[u] - {hw.Node("genre", "g")}, key=repr), 2):    # two edges per user
            val.append((u, it, 1.0))
    ordered = sorted(set(items), key=repr)
    random.seed(9)
    got = hw.validation._sample_negatives(val, interactions, ordered)
    random.seed(9)
    want = {}
    for u, i, _ in val:
        want[u] = (i, random.sample(sorted(set(items) - interactions[u], key=repr), 100))
    state_after = random.getstate()
    assert set(got) == set(want)
    for u in want:
        assert got[u][0] == want[u][0]
        assert [ordered[j] for j in got[u][1]] == want[u][1]
    random.seed(9)
    hw.validation._sample_negatives(val, interactions, ordered)
    assert random.getstate() == state_after            # the stream is consumed exactly like the restatement's


def test_ncf_eval_host_logic_with_a_stub_model():
    """The host side of validation.ncf_eval (1 positive + 100 sampled negatives per user, rank of the positive ->
    HR@10 / NDCG@10, hwer/validation.py:68-97) with the device scorer replaced by a table lookup."""
    import random
    import math
    import torch
    import hwer_b200 as hw
    users = [hw.Node("user", i) for i in range(5)]
    items = [hw.Node("item", i) for i in range(150)]
    index = {n: j for j, n in enumerate(users + items)}

    class Stub:
        def _rows_of(self, nodes):
            return torch.tensor([index[n] for n in nodes], dtype=torch.int64)

        def predict_rows(self, src, dst):
            # the user's positive (item id == user id) scores 0.9 for even users, 0.0 for odd ones; negatives 0.5
            item = dst - len(users)
            pos = item == src
            return torch.where(pos, torch.where(src % 2 == 0, 0.9, 0.0), 0.5).float()

    train = [hw.Edge(u, items[100 + j], 1.0) for j, u in enumerate(users)]
    val = [hw.Edge(u, items[j], 1.0) for j, u in enumerate(users)]
    random.seed(0)
    s = hw.validation.ncf_eval_scores(Stub(), train, val, items).numpy()
    assert s.shape == (5, 101)
    np.testing.assert_array_equal(s[:, 0], np.float32([0.9, 0.0, 0.9, 0.0, 0.9]))    # the positive is column 0
    assert (s[:, 1:] == 0.5).all()                                                    # 100 negatives, never the positive
    rank = (s[:, 1:] > s[:, :1]).sum(1)                 # what hwer_hit_rank_metrics counts on the device
    assert list(rank) == [0, 100, 0, 100, 0]
    with pytest.raises(RuntimeError):                   # the metric arithmetic has no CPU path
        hw.validation.ncf_eval(Stub(), train, val, items)
    assert hw.validation.ncf_eval_scores(Stub(), train, [], items) is None


def test_node_hash_is_recomputed_after_pickling_into_another_process():
    """Node caches hash((type, id)); string hashes differ between processes, so the cache must not travel."""
    import pickle
    import subprocess
    import hwer_b200 as hw
    blob = pickle.dumps([hw.Node("user", 3), hw.Node("item", "abc")])
    code = ("import sys, pickle; sys.path.insert(0, %r); import hwer_b200 as hw; "
            "a, b = pickle.loads(sys.stdin.buffer.read()); "
            "assert a == hw.Node('user', '3') and hash(a) == hash(hw.Node('user', 3)); "
            "assert {hw.Node('item', 'abc'): 1}[b] == 1; print('ok')" % ROOT)
    for seed in ("1", "2"):
        r = subprocess.run([sys.executable, "-c", code], input=blob, capture_output=True,
                           env=dict(os.environ, PYTHONHASHSEED=seed))
        assert r.returncode == 0 and b"ok" in r.stdout, r.stderr.decode()[-500:]


def test_sample_neighbours_shape_and_determinism():
    """Host stand-in for the reference's NeighborSampler (hwer/gcn_ncf.py:262-272): at most `fanout` distinct
    in-neighbours over the undirected edge list + one self loop per node, one list per block, seeded."""
    import hwer_b200 as hw
    rs = np.random.RandomState(3)
    n = 500
    src, dst = rs.randint(0, n, 3000), rs.randint(0, n, 3000)
    a = hw.ops.sample_neighbours(n, src, dst, fanout=2, seed=7, blocks=3)
    b = hw.ops.sample_neighbours(n, src, dst, fanout=2, seed=7, blocks=3)
    assert len(a) == 3
    nbrs = [set() for _ in range(n)]
    for s_, d_ in zip(src, dst):
        nbrs[d_].add(int(s_))
        nbrs[s_].add(int(d_))
    for (p, i), (p2, i2) in zip(a, b):
        assert np.array_equal(p, p2) and np.array_equal(i, i2)
        assert p.shape == (n + 1,) and p[0] == 0 and p[-1] == i.shape[0]
        for v in range(n):
            row = i[p[v]:p[v + 1]].tolist()
            assert row[-1] == v and 1 <= len(row) <= 3            # the self loop closes every list
            assert all(u in nbrs[v] for u in row[:-1])
            assert len(row) - 1 == min(2, sum(1 for s_, d_ in zip(src, dst) if d_ == v) + sum(1 for s_, d_ in zip(src, dst) if s_ == v))
    assert any(not np.array_equal(a[0][1], a[k][1]) for k in (1, 2))    # blocks draw independently
