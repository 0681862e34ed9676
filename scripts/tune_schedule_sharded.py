"""Round-schedule sweep for hwer_topk_sharded at N GPUs (run under torchrun): C4 catalogue (10 M x 128 split over the
ranks, top-100), whole steps timed on the device (max over ranks) for a grid of (first dense round rows, growth)
through the HWER_FIRST_ROWS / HWER_GROWTH knobs (read when an index is created; every rank walks the same grid).
Usage: torchrun --nproc-per-node N scripts/tune_schedule_sharded.py [--batches 4096,64] [--grid 4096:16,1024:32,...]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import hwer_b200 as hw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--batches", default="4096")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--grid", default="0:0,4096:16,2048:16,2048:32,1024:16,1024:32,1024:64,512:32,512:64")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    b, e = hw.sharded.partition(a.items, world, rank)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    table = torch.empty((e - b, a.dim), dtype=torch.float32, device=dev)
    for s in range(0, e - b, 2_000_000):
        t = min(e - b, s + 2_000_000)
        table[s:t] = hw.ops.unit_length(torch.randn((t - s, a.dim), generator=g, device=dev))
    shadow = hw.ops.make_shadow(table)
    out = {}
    for B in [int(x) for x in a.batches.split(",")]:
        gq = torch.Generator(device=dev).manual_seed(5)
        q = hw.ops.unit_length(torch.randn((B, a.dim), generator=gq, device=dev))
        ref = None
        for cell in a.grid.split(","):
            first, growth = [int(x) for x in cell.split(":")]
            for v, val in (("HWER_FIRST_ROWS", first), ("HWER_GROWTH", growth)):
                if val:
                    os.environ[v] = str(val)
                else:
                    os.environ.pop(v, None)
            sh = hw.sharded.ShardedTopK(table, b, shadow=shadow, max_norm=1.0001, exchange="p2p")
            key = "B%d first=%s g=%s" % (B, first or "default", growth or "default")
            try:
                for _ in range(3):
                    idx, sc = sh.topk(q, a.k)
                chk = int(idx.sum().item())
                ref = chk if ref is None else ref
                assert chk == ref, "schedule changed the answer"
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps):
                    sh.topk_p2p_async(q, a.k)
                e1.record()
                torch.cuda.synchronize()
                rc, need = sh.index.finish()
                t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sh.index.profile(True)
                for _ in range(a.steps):
                    sh.topk_p2p_async(q, a.k)
                torch.cuda.synchronize()
                st = {k_: round(v_ / a.steps, 4) for k_, v_ in sh.index.profile_stages().items()}
                sh.index.profile_read()
                sh.index.profile(False)
                out[key] = {"ms": round(float(t.item()), 4), "stages": st} if rc == 0 else "overflow(%d)" % need
            except Exception as ex:
                out[key] = "error: %s" % str(ex)[:100]
            if rank == 0:
                print(key, json.dumps(out[key]), flush=True)
            sh.close()
    if rank == 0:
        print("RESULT " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
