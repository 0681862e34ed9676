"""CPU: the on-disk table format (hwer_b200/table_io.py): round trips, shard ranges, header
validation.  The GPU leg (load straight to the device, then serve from it) is in tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

from hwer_b200 import table_io


def test_round_trip_float32_and_ranges(tmp_path):
    rs = np.random.RandomState(0)
    t = rs.standard_normal((1000, 48)).astype(np.float32)
    p = os.path.join(str(tmp_path), "t.hwer")
    table_io.save_table(p, t, node_types={"user": (0, 300), "item": (300, 1000)}, meta={"alpha": 0.5}, chunk_rows=128)
    hdr, off = table_io.read_header(p)
    assert off % 4096 == 0 and hdr["rows"] == 1000 and hdr["dim"] == 48 and hdr["dtype"] == "float32"
    assert hdr["node_types"] == {"user": [0, 300], "item": [300, 1000]} and hdr["meta"] == {"alpha": 0.5}
    full, _ = table_io.load_table(p, chunk_rows=333)
    np.testing.assert_array_equal(full.numpy(), t)
    shard, _ = table_io.load_table(p, rows=(300, 1000), chunk_rows=100)
    np.testing.assert_array_equal(shard.numpy(), t[300:])
    empty, _ = table_io.load_table(p, rows=(10, 10))
    assert empty.shape == (0, 48)


def test_round_trip_bfloat16(tmp_path):
    t = torch.randn((257, 64), generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    p = os.path.join(str(tmp_path), "s.hwer")
    table_io.save_table(p, t, unit_norm=False)
    back, hdr = table_io.load_table(p)
    assert hdr["dtype"] == "bfloat16" and not hdr["unit_norm"]
    assert torch.equal(back.view(torch.int16), t.view(torch.int16))


def test_rejects_bad_files(tmp_path):
    p = os.path.join(str(tmp_path), "bad.hwer")
    open(p, "wb").write(b"not a table")
    with pytest.raises(ValueError):
        table_io.read_header(p)
    t = np.zeros((10, 4), np.float32)
    table_io.save_table(p, t)
    data = open(p, "rb").read()
    open(p, "wb").write(data[:-8])                       # truncated payload
    with pytest.raises(ValueError):
        table_io.read_header(p)
    with pytest.raises(ValueError):
        table_io.save_table(p, t, node_types={"x": (0, 11)})
    with pytest.raises(TypeError):
        table_io.save_table(p, torch.zeros((2, 2), dtype=torch.float64))
    table_io.save_table(p, t)
    with pytest.raises(ValueError):
        table_io.load_table(p, rows=(5, 20))
