"""Where the time of one find_closest_neighbours_batch call goes on the C4 shape (10M items, 4096 user anchors):
every stage bracketed by a device synchronisation.  Usage: python scripts/time_api.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hwer_b200 as hw  # noqa: E402
from hwer_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
n, d, B, k = 10_000_000, 128, 4096, 100
g = torch.Generator(device=dev).manual_seed(0)
table = torch.empty((B + n, d), dtype=torch.float32, device=dev)
for b in range(0, B + n, 2_000_000):
    e = min(B + n, b + 2_000_000)
    table[b:e] = ops.unit_length(torch.randn((e - b, d), generator=g, device=dev))
users = [hw.Node("user", i) for i in range(B)]
model = hw.ContentRecommendation(None, {"user", "item"}, n_dims=d)
model.add_nodes(users)
model.add_node_range("item", n)
model.__build_knn__(table)
model.fit_done = True


def t(label, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    print("%-42s %8.3f ms" % (label, (time.perf_counter() - t0) / reps * 1e3), flush=True)
    return out


rows_list = t("nodes_to_idx.rows_of(4096 Nodes)", lambda: model.nodes_to_idx.rows_of(users))
arows = t("torch.tensor(rows) -> device", lambda: torch.tensor(rows_list, dtype=torch.int64, device=dev))
q = t("compose_queries", lambda: ops.compose_queries(model.device_vectors, arows, None, None))
res = t("knn.query_batch (search + finish)", lambda: model.knn.query_batch(q, "item", k=k))
t("index.topk_async only + sync", lambda: model.knn.knn["item"].topk_async(q, k, "exact", 4096, 0, want_f64=True))
out = t("rerank pair", lambda: ops.rerank(model.device_vectors, res[0], "pair", anchor_rows=arows))
t("rerank dist", lambda: ops.rerank(model.device_vectors, res[0], "dist", queries=q))
t("find_closest_neighbours_batch(Nodes)", lambda: model.find_closest_neighbours_batch("item", users, k=k))
t("find_closest_neighbours_batch(row tensor)", lambda: model.find_closest_neighbours_batch("item", arows, k=k))
ih = torch.empty((B, k), dtype=torch.int64).pin_memory()
sh = torch.empty((B, k), dtype=torch.float64).pin_memory()
t("D2H rows + scores", lambda: (ih.copy_(out[0], non_blocking=True), sh.copy_(out[1], non_blocking=True)))
