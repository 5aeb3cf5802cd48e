"""DRAM fetch granularity of random reads on this GPU: gathers rows of 16 / 32 / 64 / 128 bytes at random row indices from a
4 GB buffer (>> L2) with torch.index_select; run under
    ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:index ...
and compare the DRAM bytes with rows x row bytes."""
import sys

import torch

torch.manual_seed(0)
buf = torch.randint(0, 255, (1 << 32,), dtype=torch.uint8, device="cuda")
n = 1 << 22
for row_bytes in (16, 32, 64, 128):
    view = buf.view(-1, row_bytes // 4).view(torch.int32) if False else buf.view(torch.int32).view(-1, row_bytes // 4)
    idx = torch.randint(0, view.shape[0], (n,), device="cuda")
    torch.cuda.synchronize()
    out = torch.index_select(view, 0, idx)
    torch.cuda.synchronize()
    print(row_bytes, "rows", n, "requested MB", n * row_bytes / 1e6, file=sys.stderr)
