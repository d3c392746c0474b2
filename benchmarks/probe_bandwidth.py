#!/usr/bin/env python
"""Measured ceilings for read-only, write-only and copy streams on this GPU (torch library kernels,
used only as a yardstick for write-heavy kernels; the roofline denominator stays MEASURED_PEAKS.json)."""
import json

import torch


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    return min(e[i].elapsed_time(e[i + 1]) for i in range(reps))


def main():
    n = 1 << 28                                   # 1 GiB of fp32
    a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    b = torch.empty_like(a)
    gb = n * 4 / 1e9
    out = {
        "write_only_fill_GBps": gb / (timed(lambda: b.fill_(1.0)) * 1e-3),
        "read_only_sum_GBps": gb / (timed(lambda: a.sum()) * 1e-3),
        "copy_read_plus_write_GBps": 2 * gb / (timed(lambda: b.copy_(a)) * 1e-3),
        "one_read_two_writes_GBps": None,
    }
    c = torch.empty_like(a)
    # 1 read : 2 writes, the mix of the fused cut+pad+mix kernel on PhysioNet-shaped cycles
    def rw2():
        b.copy_(a)
        c.fill_(0.0)
    out["copy_plus_fill_sequential_GBps"] = 3 * gb / (timed(rw2) * 1e-3)
    # PCIe: one direction at a time and both at once (pinned host memory, 164 MB like one bench batch)
    m = 4096 * 4 * 2500
    h_in = torch.empty(m, dtype=torch.float32).pin_memory()
    h_out = torch.empty(m, dtype=torch.float32).pin_memory()
    d_in = torch.empty(m, dtype=torch.float32, device="cuda")
    d_out = torch.empty(m, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    mb = m * 4 / 1e9

    def h2d():
        d_in.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_out, non_blocking=True)

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)
    out["pcie_h2d_GBps"] = mb / (timed(h2d, 10) * 1e-3)
    out["pcie_d2h_GBps"] = mb / (timed(d2h, 10) * 1e-3)
    t_both = timed(both, 10)
    out["pcie_both_directions_ms_per_164MB_each"] = t_both
    out["pcie_both_directions_GBps_each"] = mb / (t_both * 1e-3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
