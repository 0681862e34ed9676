"""On-disk embedding tables for the serving path (the reference keeps its tables in
memory only and pickles whole recommender objects, so there is no reference format to mirror).

One file = one [N, d] table, row-major, little-endian:

    bytes 0..7    magic  b"HWERTBL1"
    bytes 8..15   uint64 length L of the JSON header that follows
    L bytes       JSON: {"dtype": "float32" | "bfloat16", "rows": N, "dim": d, "unit_norm": bool,
                         "node_types": {type: [first_row, end_row), ...}, "meta": {...}}
    padding       zeros up to the next multiple of 4096 (so the payload can be mapped / read with O_DIRECT)
    payload       N * d elements

`node_types` are the contiguous row ranges the per-type indexes of MultiKNN (hwer/recommendation_base.py:65-76)
are built over.  Loading streams the payload through a pinned staging buffer in row chunks, so a table larger than
host memory headroom (config C5: 16 GB of bf16 per GPU) never needs a second full host copy; a shard can load only
its own row range.
"""
import json
import os
import struct
from typing import Dict, Optional, Tuple

import numpy as np
import torch

MAGIC = b"HWERTBL1"
ALIGN = 4096
_DTYPES = {"float32": (np.float32, torch.float32, 4), "bfloat16": (np.uint16, torch.bfloat16, 2)}


def _header_bytes(header: dict) -> bytes:
    js = json.dumps(header, sort_keys=True).encode()
    head = MAGIC + struct.pack("<Q", len(js)) + js
    pad = (-len(head)) % ALIGN
    return head + b"\0" * pad


def save_table(path: str, table, node_types: Optional[Dict[str, Tuple[int, int]]] = None, unit_norm: bool = True,
               meta: Optional[dict] = None, chunk_rows: int = 1 << 18) -> None:
    """Writes a torch (CPU or CUDA; float32 or bfloat16) or numpy float32 table."""
    if isinstance(table, np.ndarray):
        table = torch.from_numpy(np.ascontiguousarray(table))
    if table.dim() != 2:
        raise ValueError("table must be [N, d]")
    name = {torch.float32: "float32", torch.bfloat16: "bfloat16"}.get(table.dtype)
    if name is None:
        raise TypeError("table must be float32 or bfloat16, got %s" % table.dtype)
    n, d = table.shape
    for t, (b, e) in (node_types or {}).items():
        if not (0 <= b <= e <= n):
            raise ValueError("node type %r has rows [%d, %d) outside the table" % (t, b, e))
    header = {"dtype": name, "rows": int(n), "dim": int(d), "unit_norm": bool(unit_norm),
              "node_types": {t: [int(b), int(e)] for t, (b, e) in (node_types or {}).items()}, "meta": meta or {}}
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(_header_bytes(header))
        for r0 in range(0, n, chunk_rows):
            blk = table[r0:r0 + chunk_rows].contiguous().cpu()
            if name == "bfloat16":
                blk = blk.view(torch.int16)
            f.write(blk.numpy().tobytes())
    os.replace(tmp, path)


def read_header(path: str) -> Tuple[dict, int]:
    """(header dict, payload offset)."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s is not a hwer_b200 table file" % path)
        (ln,) = struct.unpack("<Q", f.read(8))
        if ln > (1 << 26):
            raise ValueError("%s: corrupt header" % path)
        header = json.loads(f.read(ln).decode())
    off = 16 + ln
    off += (-off) % ALIGN
    name = header.get("dtype")
    if name not in _DTYPES:
        raise ValueError("%s: unknown dtype %r" % (path, name))
    want = off + header["rows"] * header["dim"] * _DTYPES[name][2]
    if os.path.getsize(path) < want:
        raise ValueError("%s: truncated (need %d bytes)" % (path, want))
    return header, off


def load_table(path: str, device="cpu", rows: Optional[Tuple[int, int]] = None, chunk_rows: int = 1 << 18):
    """Returns (tensor [rows, d] on `device`, header).  `rows` = (first, end) loads one shard's range only."""
    header, off = read_header(path)
    np_dt, th_dt, esz = _DTYPES[header["dtype"]]
    n, d = header["rows"], header["dim"]
    b, e = (0, n) if rows is None else rows
    if not (0 <= b <= e <= n):
        raise ValueError("rows [%d, %d) outside the table of %d rows" % (b, e, n))
    device = torch.device(device)
    out = torch.empty((e - b, d), dtype=th_dt, device=device)
    mm = np.memmap(path, dtype=np_dt, mode="r", offset=off, shape=(n, d))
    stage = None
    if device.type == "cuda":
        stage = [torch.empty((min(chunk_rows, max(e - b, 1)), d), dtype=th_dt).pin_memory() for _ in range(2)]
        events = [None, None]
    for i, r0 in enumerate(range(b, e, chunk_rows)):
        r1 = min(e, r0 + chunk_rows)
        src = torch.from_numpy(np.array(mm[r0:r1]))         # a writable copy of the chunk
        if header["dtype"] == "bfloat16":
            src = src.view(torch.bfloat16)
        if stage is None:
            out[r0 - b:r1 - b] = src
        else:
            s = stage[i & 1]
            if events[i & 1] is not None:
                events[i & 1].synchronize()          # the copy that last used this staging buffer has finished
            s[:r1 - r0].copy_(src)
            out[r0 - b:r1 - b].copy_(s[:r1 - r0], non_blocking=True)
            events[i & 1] = torch.cuda.Event()
            events[i & 1].record()
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    del mm
    return out, header
