#!/usr/bin/env python
"""Benchmark of the serving hot path on BASELINE.json's headline workload (config C4):
exact top-100 by cosine over a synthetic 10M x 128 unit-norm catalogue, item-sharded over N GPUs.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on the host cores
  python bench.py --workload c3|c5|c2 ...                  # the other BASELINE.json configs (see run_c3 / run_c2;
                                                           # c5 = the C4 code path at 62.5M rows per GPU, top-1000)

One JSON line on stdout (rank 0).  A "step" is one batch of `--batch` queries answered against the whole
catalogue.  `value` = queries/s with queries and catalogue resident in HBM; `e2e` = the same through the public
API with pinned-host queries copied in and results copied out every step.  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "top-100 cosine queries/sec @10Mx128 items"
METRIC_C5 = "top-1000 cosine queries/sec @500Mx128 items (62.5M rows per GPU)"
METRIC_C3 = "top-100 cosine queries/sec, all 138,493 users x 27,278 items x d=256"
UNIT = "queries/s"
CHUNK = 1_000_000          # rows generated per seeded chunk: the table is the same for every shard count


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c5", "c3", "c2"])
    ap.add_argument("--items", type=int, default=None)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--k", type=int, default=None)
    ap.add_argument("--no-extras", action="store_true", help="c4: skip the c3 / c2 summary blocks of the default line")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--mode", default="exact", choices=["exact", "bf16"])
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--sweep", default="1,64", help="extra batch sizes reported under 'sweep' (N=1 only)")
    ap.add_argument("--cpu-sample-rows", type=int, default=200_000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: result exchange over NVLink peer stores fused into the final kernel, or NCCL all-gather")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.workload == "c5":        # configs[4]: 500M x 128 over 8 GPUs = 62.5M rows per GPU, top-1000 (weak scaling in N)
        a.items = a.items or 62_500_000 * world
        a.k = a.k or 1000
    else:
        a.items = a.items or 10_000_000
        a.k = a.k or 100
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tc_burst=float(d["bf16_tflops"]),
                    tc_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), basis="of measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, basis="of fallback")


def ncu_traffic(batch):
    """DRAM bytes (read + write) of the filter kernel per step, from the newest committed ncu capture
    (profiles/r*_traffic.json, scripts/gpu_traffic.sh); None if this batch size was not captured."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None, None
    try:
        b = json.load(open(files[-1]))["batches"].get(str(batch))
        return b["dram_read_bytes_per_step"] + b["dram_write_bytes_per_step"], os.path.basename(files[-1])
    except Exception:
        return None, None


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = "/tmp/hwer_clocks_%d_%d.csv" % (os.getpid(), gpu_index)
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(gpu_index)], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------- CPU arms
def cpu_table(rows, d, seed=0):
    rs = np.random.RandomState(seed)
    c = rs.standard_normal((rows, d)).astype(np.float32)
    g = rs.standard_normal((rows, d)).astype(np.float32)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hwer_oracle as O
    return O.blend_normalize(c, g, 0.5)


def build_cpu_model(table, queries):
    """The reference's index + anchors (restated in oracle/hwer_oracle.py): one KD-tree per node type."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hwer_oracle as O
    users = [O.Node("user", i) for i in range(len(queries))]
    items = [O.Node("item", i) for i in range(len(table))]
    m = O.OracleRecommender({"user", "item"}, n_dims=table.shape[1])
    m.add_nodes(users + items)
    t0 = time.time()
    m.build_knn(np.concatenate([queries, table]).astype(np.float32))
    return m, users, time.time() - t0


_CPU_MODEL = None      # set before forking so pool workers inherit the built trees instead of unpickling them


def _cpu_loop(args):
    model, anchors, k = args
    if model is None:
        model = _CPU_MODEL
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hwer_oracle as O
    O.model_get_topk_knn(model, anchors, "item", k=k)     # hwer/validation.py:30-35: serial per-anchor loop
    return len(anchors)


def cpu_baseline(table_sample, queries, k, full_rows, seconds):
    """Single core (the reference path is serial by construction): queries/s on the sample, scaled by
    sample_rows / catalogue_rows (a KD-tree in d = 128 degenerates to a linear scan)."""
    model, users, build_s = build_cpu_model(table_sample, queries)
    _cpu_loop((model, users[:2], k))
    n, t0 = 0, time.time()
    while time.time() - t0 < seconds and n < len(users):
        n += _cpu_loop((model, users[n:n + 4], k))
    dt = time.time() - t0
    qps_sample = n / dt
    scale = len(table_sample) / float(full_rows)
    # a best-effort CPU restatement that is NOT the reference's algorithm -- fp32 BLAS `Q @ V.T` over the
    # same sample with every host core + argpartition -- reported separately and labelled, also scaled by rows
    blas = None
    try:
        q32 = np.ascontiguousarray(queries[:64], dtype=np.float32)
        v32 = np.ascontiguousarray(table_sample, dtype=np.float32)
        (q32 @ v32.T)
        reps, tb = 0, time.time()
        while time.time() - tb < min(3.0, seconds) or reps < 2:
            sc = q32 @ v32.T
            part = np.argpartition(-sc, k, axis=1)[:, :k]
            reps += 1
        db = time.time() - tb
        blas = {"value": reps * q32.shape[0] / db * scale, "unit": UNIT, "cores": os.cpu_count(), "kind": "restatement",
                "sample": "numpy fp32 BLAS Q @ V.T + argpartition (not the reference's KD-tree), %d queries x %d rows, "
                          "%d passes in %.2f s, scaled by rows" % (q32.shape[0], len(v32), reps, db)}
        del sc, part
    except Exception as ex:                      # a reported extra, never a reason to lose the line
        blas = {"error": str(ex)[:120]}
    return {"value": qps_sample * scale, "unit": UNIT, "cores": 1, "kind": "port", "blas_all_cores": blas,
            "sample": "oracle/hwer_oracle.py restatement of find_closest_neighbours (sklearn KDTree, float64) on the "
                      "first %d rows of the catalogue, %d queries in %.1f s = %.2f q/s on the sample, scaled by "
                      "%d/%d rows; KD-tree build %.1f s not included" % (len(table_sample), n, dt, qps_sample,
                                                                        len(table_sample), full_rows, build_s)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    rows = min(a.cpu_sample_rows, a.items)
    cores = os.cpu_count() or 1
    table = cpu_table(rows, a.dim)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hwer_oracle as O
    per_step = max(cores * 2, 16)
    rs = np.random.RandomState(2)
    queries = O.unit_length(rs.standard_normal((per_step, a.dim)).astype(np.float32), axis=1)
    global _CPU_MODEL
    model, users, build_s = build_cpu_model(table, queries)
    _CPU_MODEL = model
    ctx = mp.get_context("fork")                    # workers share the built trees copy-on-write
    slices = [users[i::cores] for i in range(cores)]
    slices = [s for s in slices if s]
    with ctx.Pool(len(slices)) as pool:
        def step():
            return sum(pool.map(_cpu_loop, [(None, s, a.k) for s in slices]))
        for _ in range(a.warmup):
            step()
        t0 = time.time()
        n = 0
        for _ in range(a.steps):
            n += step()
        dt = time.time() - t0
    scale = rows / float(a.items)
    v = n / dt * scale
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4: 10M x 128 catalogue, exact top-%d, the reference's serial KD-tree loop run on "
                                   "every host core in parallel" % a.k, "items": a.items, "dim": a.dim, "k": a.k,
                       "queries_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": len(slices), "kind": "port",
                             "sample": "oracle port of hwer find_closest_neighbours (sklearn KDTree float64) on a %d-row "
                                       "sample, %d queries per step over %d forked workers, scaled by %d/%d rows; tree "
                                       "build %.1f s excluded" % (rows, per_step, len(slices), rows, a.items, build_s)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def make_shard(hw, torch, a, begin, end, dev):
    """Rows [begin, end) of the global table V = unit(alpha unit(C) + (1-alpha) unit(G)); chunk c of CHUNK rows is
    generated from seeds (c, 10^6 + c), so the catalogue does not depend on the shard count.  The blend+normalise
    kernel is timed on the whole shard (best of three launches after one untimed call)."""
    rows = end - begin
    content = torch.empty((rows, a.dim), dtype=torch.float32, device=dev)
    collab = torch.empty((rows, a.dim), dtype=torch.float32, device=dev)
    for c in range(begin // CHUNK, (end + CHUNK - 1) // CHUNK):
        cb, ce = c * CHUNK, min((c + 1) * CHUNK, a.items)
        g1 = torch.Generator(device=dev).manual_seed(c)
        g2 = torch.Generator(device=dev).manual_seed(1_000_000 + c)
        cc = torch.randn((ce - cb, a.dim), generator=g1, device=dev)
        gg = torch.randn((ce - cb, a.dim), generator=g2, device=dev)
        lo, hi = max(cb, begin), min(ce, end)
        content[lo - begin:hi - begin] = cc[lo - cb:hi - cb]
        collab[lo - begin:hi - begin] = gg[lo - cb:hi - cb]
        del cc, gg
    table, shadow = hw.ops.blend_normalize(content, collab, a.alpha)
    torch.cuda.synchronize()
    blend_ms = float("inf")
    for _ in range(3):                               # best of three whole-shard launches, like MEASURED_PEAKS' copy figure
        del table, shadow
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        table, shadow = hw.ops.blend_normalize(content, collab, a.alpha)
        e1.record()
        torch.cuda.synchronize()
        blend_ms = min(blend_ms, e0.elapsed_time(e1))
    blend_bytes = rows * a.dim * (4 + 4 + 4) + rows * shadow.shape[1] * 2
    del content, collab
    return table, shadow, blend_ms, blend_bytes


def run_b200(a):
    import torch
    import torch.distributed as dist
    import hwer_b200 as hw
    from hwer_b200 import _native
    _native.lib()                                    # fail loudly now if the CUDA library is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.workload in ("c3", "c2"):
        if a.workload == "c3":
            line = run_c3(a, hw, torch, dist, dev, world, rank)
        else:
            c2 = run_c2(a, hw, torch, dev) if rank == 0 else None
            line = c2 and {"metric": "validation.extraction_efficiency seconds, ML-1M shape", "value": c2["seconds"],
                           "unit": "s", "n_gpus": 1, "steps": 1, "warmup": 1, "ms_per_step": c2["seconds"] * 1e3,
                           "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64 metrics",
                           "data": "synthetic", "config": {"workload": c2["workload"]}, "c2": c2}
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    pk = peaks()
    begin, end = hw.sharded.partition(a.items, world, rank)
    table, shadow, blend_ms, blend_bytes = make_shard(hw, torch, a, begin, end, dev)
    viol, _, _, _, max_norm = hw.ops.norm_stats(table)
    assert viol == 0
    shard = hw.sharded.ShardedTopK(table, begin, shadow=shadow, max_norm=max_norm, exchange=a.exchange)
    index = shard.index
    d_pad = shadow.shape[1]

    def queries_for(B):
        g = torch.Generator(device=dev).manual_seed(2)
        return hw.ops.unit_length(torch.randn((B, a.dim), generator=g, device=dev))

    def step_device(q):
        if world > 1 and a.exchange == "p2p":
            idx, _, s64 = shard.topk_p2p_async(q, a.k, a.mode, want_f64=True)
            return idx, s64
        idx, _, s64 = index.topk_async(q, min(a.k, index.n), a.mode, idx_offset=begin, want_f64=True)
        if world > 1:
            idx, s64 = hw.sharded.pad_local_result(idx, s64, a.k)
            gs, gi = hw.sharded.gather_shard_results(idx, s64)
            return hw.ops.merge_topk(gs, gi)
        return idx, s64

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stages = {}                      # per-step ms of the last profiled pass: filter / select / final / exchange

    def measure(B, steps, warmup, profile):
        q = queries_for(B)
        for _ in range(warmup):
            step_device(q)
        rc, need = index.finish()
        if rc != 0:
            raise RuntimeError("candidate overflow in warm-up (needed cap %d)" % need)
        index.profile(profile)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = step_device(q)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        rc, need = index.finish()
        if rc != 0:
            raise RuntimeError("candidate overflow in the timed region (needed cap %d)" % need)
        stages.clear()
        if profile:
            stages.update({k_: v_ / steps for k_, v_ in index.profile_stages().items()})
        filt_ms, filt_l, other_l = index.profile_read() if profile else (0.0, 0, 0)
        index.profile(False)
        return ms, filt_ms, filt_l, other_l, out

    def roofline(B, filt_ms_per_step):
        rows = end - begin
        byt = rows * d_pad * 2.0 + B * a.dim * 4.0 + B * a.k * 8.0
        # measured DRAM traffic of the same kernel (ncu, per step): only valid for the captured shape (N = 1, C4)
        traffic, src = ncu_traffic(B) if (world == 1 and a.items == 10_000_000 and a.dim == 128 and a.k == 100) else (None, None)
        extra = {"traffic": traffic, "traffic_unit": "bytes per step, dram read+write of all filter launches (ncu)",
                 "traffic_source": src, "algorithmic_bytes": byt, "kernel": "score_filter_tc_kernel (+ spill_extract_kernel behind it)",
                 "ms_per_step": filt_ms_per_step}
        if B >= 256:
            flops = 2.0 * B * rows * a.dim
            ach = flops / (filt_ms_per_step * 1e-3) / 1e12
            # the timed region of the default run is a fraction of a second: the burst cuBLAS figure is the basis
            # (the sustained, power-capped one is reported next to it)
            return dict({"bound": "tensor", "achieved": ach, "peak": pk["tc_burst"], "unit": "TFLOP/s",
                         "frac": ach / pk["tc_burst"], "basis": pk["basis"] + " (burst bf16 cuBLAS)",
                         "peak_sustained": pk["tc_sustained"], "frac_of_sustained": ach / pk["tc_sustained"],
                         "algorithmic_flops": flops}, **extra)
        ach = byt / (filt_ms_per_step * 1e-3) / 1e9
        return dict({"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                     "basis": pk["basis"] + " (copy bandwidth)"}, **extra)

    # ---- multi-GPU preflight: the peer-memory exchange must reproduce the NCCL all-gather + merge answer bit for
    # bit before it is timed; if a peer fails to show up (HWER_E_PEER) or the answers differ, every rank switches to
    # the NCCL exchange together and the JSON line says so (the number is then the plain path's, not a crash)
    exchange_note = None
    if world > 1 and a.exchange == "p2p":
        qp = queries_for(a.batch)
        ok, why = 1, ""
        try:
            idx_p, sc_p = shard.topk(qp, a.k, a.mode)
        except _native.HwerError as ex:
            ok, why = 0, str(ex)
        agree = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        plain = hw.sharded.ShardedTopK(table, begin, shadow=shadow, max_norm=max_norm, exchange="nccl")
        if int(agree.item()) == 1:
            idx_n, sc_n = plain.topk(qp, a.k, a.mode)
            same = int(torch.equal(idx_p, idx_n) and torch.equal(sc_p, sc_n))
            agree = torch.tensor([same], dtype=torch.int32, device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            why = why or "peer exchange and NCCL exchange disagreed in the preflight step"
        if int(agree.item()) == 0:
            exchange_note = "fell back to the NCCL exchange: %s" % (why or "a peer rank reported a failure")
            a.exchange = "nccl"
            shard.close()
            shard, index = plain, plain.index
        else:
            plain.close()
            del plain

    # ---- headline: device-resident throughput (events on the launching stream), clocks sampled meanwhile
    sampler = ClockSampler(local) if rank == 0 else None
    ms, _, _, _, out = measure(a.batch, a.steps, a.warmup, profile=False)
    # same steps again with the filter kernel bracketed by events: kernel time for the roofline + launch counts
    # (the clock sampler keeps running: both passes are the same work, and one pass can be shorter than a sample)
    pms, filt_ms, filt_l, other_l, _ = measure(a.batch, a.steps, 1, profile=True)
    clocks = sampler.stop() if sampler else None
    launches_per_step = (filt_l + other_l) / a.steps + (1 if world > 1 and a.exchange == "nccl" else 0)
    roof = roofline(a.batch, filt_ms / a.steps)
    roof["share_of_step"] = (filt_ms / a.steps) / (pms / a.steps)
    stage_ms = dict(stages, step=pms / a.steps)
    # ---- end to end.  N = 1: through the reference-facing API, model.find_closest_neighbours_batch(node_type, anchors,
    # k) -- Node -> row lookup on the host, anchor rows copied in, query composition, search, score convention +
    # per-anchor ordering, rows and scores copied out to pinned host memory, every step.  The 10M items are a node
    # RANGE of the model (no Python object per item); the anchors are real Node objects.
    # N > 1: ShardedTopK.topk with pinned-host query vectors in and [B, k] results out on every rank (the hwer model
    # classes serve one GPU; the sharded index is the multi-GPU entry point).
    idx_host = torch.empty((a.batch, a.k), dtype=torch.int64).pin_memory()
    e2e_api = None
    if world == 1 and a.workload == "c4":
        users = [hw.Node("user", i) for i in range(a.batch)]
        table_all = torch.cat([queries_for(a.batch), table])
        shadow_all = torch.cat([hw.ops.make_shadow(table_all[:a.batch]), shadow])
        shard = index = table = shadow = None          # the model below owns the (one) copy of the tables
        torch.cuda.empty_cache()
        model = hw.ContentRecommendation(None, {"user", "item"}, n_dims=a.dim, mode=a.mode)
        model.add_nodes(users)
        model.add_node_range("item", a.items)
        model.__build_knn__(table_all, shadow=shadow_all)
        model.fit_done = True
        index = model.knn.knn["item"]
        table, shadow = table_all[a.batch:], shadow_all[a.batch:]
        shard = None
        sc_host = torch.empty((a.batch, a.k), dtype=torch.float64).pin_memory()
        e2e_api = "ContentRecommendation.find_closest_neighbours_batch('item', <%d user Nodes>, k=%d)" % (a.batch, a.k)
        h2d, d2h = a.batch * 8, a.batch * a.k * 16

        def step_e2e():
            rows, sc = model.find_closest_neighbours_batch("item", users, k=a.k)
            idx_host.copy_(rows, non_blocking=True)
            sc_host.copy_(sc, non_blocking=True)
            torch.cuda.synchronize()
    else:
        q_host = queries_for(a.batch).cpu().pin_memory()
        # N > 1: every rank uploads the (replicated) query batch and copies out the rows of the queries it merged
        # (ShardedTopK.topk(owned=True)): each query's answer reaches exactly one host buffer, 1/N of the result per rank
        lo, hi = shard.owner_range(a.batch) if world > 1 else (0, a.batch)
        idx_host = torch.empty((hi - lo, a.k), dtype=torch.int64).pin_memory()
        sc_host = torch.empty((hi - lo, a.k), dtype=torch.float32).pin_memory()
        e2e_api = ("ShardedTopK.topk(<pinned host queries>, k=%d, owned=True): rank r copies out the %d queries it merged"
                   % (a.k, hi - lo)) if world > 1 else "TopKIndex.topk"
        h2d, d2h = a.batch * a.dim * 4, (hi - lo) * a.k * 12

        def step_e2e():
            q = q_host.to(dev, non_blocking=True)
            if world > 1:
                idx, sc = shard.topk(q, a.k, a.mode, owned=True)
            else:
                idx, sc = index.topk(q, a.k, a.mode, idx_offset=begin)
            idx_host.copy_(idx, non_blocking=True)
            sc_host.copy_(sc, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(max(1, a.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, a.steps // 2)
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    api_same = None
    if world == 1 and a.workload == "c4":
        # the API answers in global rows (users first): the same items as the device-resident pass (an anchor's row is
        # re-normalised by the API, so a k-th / (k+1)-th pair closer than ~1e-8 may swap: reported, not asserted)
        api_same = int(idx_host.sum().item()) - a.batch * a.k * a.batch == int(out[0].sum().item())

    sweep = {}
    if a.sweep:
        for B in [int(x) for x in a.sweep.split(",") if x]:
            # small batches are ~0.5 ms steps: 200 of them, so that the timed region (0.1 s) is not a clock ramp
            st = max(a.steps, 200 if B <= 256 else 20)
            sms, sf, _, _, _ = measure(B, st, a.warmup, profile=False)
            _, sf, _, _, _ = measure(B, st, 1, profile=True)
            r = roofline(B, sf / st)
            sweep[str(B)] = {"value": B * st / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms / st, "roofline": r,
                             "stage_ms": dict(stages)}

    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            rows = min(a.cpu_sample_rows, end - begin)
            cpu = cpu_baseline(table[:rows].cpu().numpy(), queries_for(64).cpu().numpy(), a.k, a.items, a.cpu_seconds)
        idx_chk = out[0]
        line = {
            "metric": METRIC if a.workload == "c4" else METRIC_C5, "value": a.batch * a.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong" if a.workload == "c4" else "weak", "vs_baseline": None, "dtype": "bf16 tensor-core filter + f64 re-score" if a.mode == "exact" else "bf16",
            "data": "synthetic",
            "config": {"workload": "%s: synthetic %s x %d unit-norm catalogue (alpha=%.2f blend of two seeded Gaussian "
                                   "tables), exact top-%d by cosine, query batch %d, item-sharded over %d GPU(s)"
                                   % (a.workload.upper(), "{:,}".format(a.items), a.dim, a.alpha, a.k, a.batch, world),
                       "items": a.items, "dim": a.dim, "k": a.k, "batch": a.batch, "mode": a.mode,
                       "parallelism": ("item-shard x%d + %s" % (world, "peer-store exchange fused into the final kernel + owner merge"
                                                                      if a.exchange == "p2p" else "NCCL all-gather + merge"))
                       if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (%.2f GB bf16 shadow per GPU streamed every step)"
                             % ((end - begin) * d_pad * 2 / 1e9)},
            "e2e": {"value": a.batch * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_steps,
                    "api": e2e_api, "same_items_as_device_pass": api_same},
            "stage_ms": stage_ms,
            "gpu_launches": int(round(launches_per_step * a.steps)),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "sweep": sweep,
            "blend_normalize": {"achieved": blend_bytes / (blend_ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                "frac": blend_bytes / (blend_ms * 1e-3) / 1e9 / pk["hbm"], "bound": "hbm"},
            "result_checksum": int(idx_chk.sum().item()),
        }
        if exchange_note:
            line["config"]["exchange_note"] = exchange_note
        if world == 1 and a.workload == "c4" and not a.no_extras:
            try:
                line["aux_kernels"] = aux_kernels(hw, torch, table, pk, dev)
            except Exception as ex:
                line["aux_kernels"] = {"error": str(ex)[:200]}
            # summary blocks of the other single-GPU configs (full lines: --workload c3 / c2)
            try:
                c3 = run_c3(a, hw, torch, dist, dev, 1, 0, steps=max(3, a.steps // 4), cpu=False)
                line["c3"] = {k_: c3[k_] for k_ in ("metric", "value", "unit", "ms_per_step", "e2e", "stage_ms", "roofline")}
            except Exception as ex:                                   # the headline line must not die with an extra
                line["c3"] = {"error": str(ex)[:200]}
            try:
                line["c2"] = run_c2(a, hw, torch, dev)
            except Exception as ex:
                line["c2"] = {"error": str(ex)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        shard.close()
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- C3: all users, small catalogue
def run_c3(a, hw, torch, dist, dev, world, rank, steps=None, cpu=True):
    """BASELINE.json configs[2]: ML-20M shape, 138,493 users x 27,278 items, d = 256, top-100 for EVERY user.  The
    27.9 MB item table is replicated; the users (queries) are sharded over the ranks, no collective on the data path.
    A step = one pass over all users.  `value`: anchor rows resident on the device; `e2e`: the hwer API with Node
    anchors, rows + scores copied to pinned host memory."""
    U, I, d, k = 138_493, 27_278, 256, 100
    steps = steps or a.steps
    g1 = torch.Generator(device=dev).manual_seed(600)
    g2 = torch.Generator(device=dev).manual_seed(601)
    table, shadow = hw.ops.blend_normalize(torch.randn((U + I, d), generator=g1, device=dev),
                                           torch.randn((U + I, d), generator=g2, device=dev), a.alpha)
    users = [hw.Node("user", i) for i in range(U)]
    items = [hw.Node("item", i) for i in range(I)]
    model = hw.ContentRecommendation(None, {"user", "item"}, n_dims=d, mode=a.mode)
    model.add_nodes(users + items)
    model.__build_knn__(table, shadow=shadow)
    model.fit_done = True
    b, e = hw.sharded.partition(U, world, rank)
    my_users = users[b:e]
    my_rows = torch.arange(b, e, dtype=torch.int64, device=dev)
    index = model.knn.knn["item"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(a.warmup, 3)):
        out = model.find_closest_neighbours_batch("item", my_rows, k=k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = model.find_closest_neighbours_batch("item", my_rows, k=k)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    index.profile(True)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.find_closest_neighbours_batch("item", my_rows, k=k)
    torch.cuda.synchronize()
    pms = (time.perf_counter() - t0) * 1e3
    stages = {k_: v_ / steps for k_, v_ in index.profile_stages().items()}
    filt_ms, filt_l, other_l = index.profile_read()
    index.profile(False)
    idx_host = torch.empty((e - b, k), dtype=torch.int64).pin_memory()
    sc_host = torch.empty((e - b, k), dtype=torch.float64).pin_memory()

    def step_e2e():
        # chunks of 32,768 anchors: chunk i's rows / scores are copied out while chunk i + 1 is searched
        model.find_closest_neighbours_batch_to_host("item", my_users, k=k, out=(idx_host, sc_host))
        torch.cuda.synchronize()

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, steps // 2)
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    pk = peaks()
    flops = 2.0 * (e - b) * I * d
    ach = flops / (filt_ms / steps * 1e-3) / 1e12 if filt_ms > 0 else 0.0
    line = {
        "metric": METRIC_C3, "value": U * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16 tensor-core filter + f64 re-score", "data": "synthetic",
        "config": {"workload": "C3: ML-20M shape, %d users x %d items, d=%d, exact top-%d for every user through "
                               "find_closest_neighbours_batch; item table replicated, users sharded over %d GPU(s), "
                               "no data-path collective" % (U, I, d, k, world),
                   "users": U, "items": I, "dim": d, "k": k, "parallelism": "query-shard x%d" % world,
                   "l2": "the 14 MB bf16 item table is L2-resident by design; every step streams %d MB of anchors and "
                         "writes %d MB of results" % ((e - b) * d * 4 // 2**20, (e - b) * k * 16 // 2**20)},
        "e2e": {"value": U * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": (e - b) * 8,
                "d2h_bytes_per_step": (e - b) * k * 16, "ms_per_step": e2e_ms / e2e_steps,
                "api": "ContentRecommendation.find_closest_neighbours_batch_to_host('item', <%d user Nodes>, k=%d): 32,768-anchor "
                       "chunks, result copies overlapped with the next chunk's search" % (e - b, k)},
        "gpu_launches": int(filt_l + other_l) + 2 * steps,
        "stage_ms": dict(stages, step=pms / steps),
        "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["tc_burst"], "unit": "TFLOP/s", "frac": ach / pk["tc_burst"],
                     "basis": pk["basis"] + " (burst bf16 cuBLAS)", "algorithmic_flops": flops,
                     "kernel": "score_filter_tc_kernel over this rank's users", "ms_per_step": filt_ms / steps,
                     "share_of_step": (filt_ms / steps) / (pms / steps), "traffic": None,
                     "note": "the contraction is %.2f TFLOP per pass: the step is bound by per-query selection, "
                             "re-scoring and result traffic, not by the tensor pipe" % (flops / 1e12)},
        "result_checksum": int(out[0].sum().item()),
    }
    if cpu and rank == 0 and not a.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import hwer_oracle as O
        t_np = table.cpu().numpy()
        m = O.OracleRecommender({"user", "item"}, n_dims=d)
        ou = [O.Node("user", i) for i in range(U)]
        m.add_nodes(ou + [O.Node("item", i) for i in range(I)])
        tb = time.time()
        m.build_knn(t_np)
        build_s = time.time() - tb
        n, t0 = 0, time.time()
        while time.time() - t0 < a.cpu_seconds:
            O.model_get_topk_knn(m, ou[n:n + 8], "item", k=k)
            n += 8
        dt = time.time() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "oracle port of validation.model_get_topk_knn (serial find_closest_neighbours, "
                                          "sklearn KDTree float64) over the full %d x %d item table for the first %d of "
                                          "%d users in %.1f s; tree build %.1f s excluded" % (I, d, n, U, dt, build_s)}
    return line


# --------------------------------------------------------------------------------------------- C2: full validation pass
def aux_kernels(hw, torch, table, pk, dev):
    """Roofline lines of the gather kernels next to the search (SURVEY 8d names HBM for them): pair scores
    (`predict`), query composition with positive / negative lists, row gathers -- random 512-byte rows of the 10 M x 128
    table, so the ceiling is HBM's random-row rate, reported against the measured copy bandwidth.  Device-timed, best
    of five launches each."""
    n, d = table.shape
    g = torch.Generator(device=dev).manual_seed(9)
    P = 1 << 22
    src = torch.randint(0, n, (P,), generator=g, device=dev)
    dst = torch.randint(0, n, (P,), generator=g, device=dev)
    B = 1 << 16
    anchors = torch.randint(0, n, (B,), generator=g, device=dev)
    ptr = torch.arange(0, 4 * B + 1, 4, dtype=torch.int64, device=dev)
    pos = (ptr, torch.randint(0, n, (4 * B,), generator=g, device=dev))
    neg = (ptr, torch.randint(0, n, (4 * B,), generator=g, device=dev))

    def best(fn):
        fn()
        torch.cuda.synchronize()
        ms = float("inf")
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = min(ms, e0.elapsed_time(e1))
        return ms

    out = {}
    for name, fn, byt in (
            ("pair_score", lambda: hw.ops.pair_score(table, src, dst), P * (2 * d * 4 + 16 + 4)),
            ("compose_queries", lambda: hw.ops.compose_queries(table, anchors, pos, neg), B * (9 * d * 4 + d * 4 + 72)),
            ("gather_rows", lambda: hw.ops.gather_rows(table, src[:1 << 20]), (1 << 20) * (2 * d * 4 + 8))):
        ms = best(fn)
        out[name] = {"ms": ms, "achieved": byt / (ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                     "frac": byt / (ms * 1e-3) / 1e9 / pk["hbm"], "bound": "hbm (random 512-byte rows)",
                     "algorithmic_bytes": byt}
    return out


def run_c2(a, hw, torch, dev):
    """BASELINE.json configs[1]: ML-1M shape (6,040 users x 3,706 items, d = 128): one validation.extraction_efficiency
    call end to end -- top-200 for every edge source, train-item filter, Recall@K / NDCG / diversity on the device,
    ncf_eval -- on the inputs of tests/golden/reference_c2.npz, whose metric values (from the reference itself) the
    result is checked against.  `reference_seconds` is what the unmodified reference took in the build container."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import random
    from conftest import synthetic_case, synthetic_edges
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_c2.npz"))
    nu, ni, d = [int(x) for x in g["shape"]]
    _, collab = synthetic_case(nu, ni, d, int(g["seeds"][0]))
    users = [hw.Node("user", i) for i in range(nu)]
    items = [hw.Node("item", i) for i in range(ni)]
    tr, vl = synthetic_edges(nu, ni, int(g["seeds"][1]))
    train = [hw.Edge(users[u], items[i], w) for u, i, w in tr]
    val = [hw.Edge(users[u], items[i], w) for u, i, w in vl]
    model = hw.GcnNCF(None, {"user", "item"}, n_dims=d)
    model.fit(users + items, train, None, collaborative_vectors=collab)
    hw.validation.extraction_efficiency(model, train[:2000], val[:200], None, "item")       # warm-up (allocations)
    random.seed(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = hw.validation.extraction_efficiency(model, train, val, None, "item")
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    want = dict(zip([str(x) for x in g["metric_keys"]], g["metric_values"]))
    err = max(abs(res["metrics"][k_] - v_) for k_, v_ in want.items())
    ref_s = 58.7
    return {"workload": "C2: ML-1M shape, %d users x %d items, d=%d: validation.extraction_efficiency end to end "
                        "(%d train / %d validation edges)" % (nu, ni, d, len(train), len(val)),
            "seconds": secs, "retrieval_seconds": res["metrics"]["retrieval_time"],
            "reference_seconds": ref_s, "reference_retrieval_seconds": float(g["retrieval_time"][0]),
            "reference_seconds_source": "the unmodified reference run by oracle/make_golden_c2c3.py in the build container "
                                        "(8 vCPU, single-threaded path)",
            "speedup_vs_reference": ref_s / secs, "max_metric_error_vs_reference": err,
            "metrics": {k_: res["metrics"][k_] for k_ in want}}


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
