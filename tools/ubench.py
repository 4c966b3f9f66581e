#!/usr/bin/env python
"""Micro-benchmarks of the sm_100a building blocks (csrc/ubench_sm100.cu) -> one table.
    python tools/ubench.py [--grid 1|148] [--iters 512] > gpurun_out/ubench.txt
Each MMA dispatch is M x N x 16 (16-bit operands); 'ideal' is 8192 dense flop/clk/SM."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "burst-attention_b200"))
import torch  # noqa: E402

from burst_attn import native  # noqa: E402

MMA = {  # mode: (label, M per SM, N, smem operand bytes per SM per dispatch)
    0: ("SS  M128 N128 cta1", 128, 128, 4096 + 4096), 1: ("TS  M128 N128 cta1", 128, 128, 4096),
    2: ("SS  M128 N256 cta1", 128, 256, 4096 + 8192), 3: ("TS  M128 N256 cta1", 128, 256, 8192),
    4: ("SS  M128 N64  cta1", 128, 64, 4096 + 2048),
    5: ("SS  M256 N128 cta2", 128, 128, 4096 + 2048), 6: ("TS  M256 N128 cta2", 128, 128, 2048),
    7: ("SS  M256 N256 cta2", 128, 256, 4096 + 4096),
}


def run(lib, mode, iters, grid):
    out = (ctypes.c_int64 * 4)()
    native.check_selftest(lib.ba_ubench(mode, iters, grid, out, native.stream_ptr()), f"ba_ubench({mode})")
    return list(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=512)
    ap.add_argument("--grid", type=int, default=1)
    ap.add_argument("--part", default="all", choices=["all", "single", "pair"],
                    help="single-CTA modes / CTA-pair modes (run them in separate processes: a trapped kernel "
                         "poisons the CUDA context)")
    a = ap.parse_args()
    torch.cuda.init()
    lib = native.selftest_lib()
    native.check(native.lib().ba_device_check(), "ba_device_check")
    print(f"# ubench grid={a.grid} iters={a.iters} (cycles are per SM, max over CTAs)")
    lat = run(lib, 12, 1, 1)[0]
    print(f"single MMA issue -> commit -> mbarrier wait: {lat} clk")
    print(f"{'mma chain':22s} {'clk/dispatch':>12s} {'ideal':>6s} {'rate':>6s} {'smem B/clk':>10s}")
    for mode, (label, m, n, smem_bytes) in MMA.items():
        if (mode >= 5) != (a.part == "pair") and a.part != "all":
            continue
        for _ in range(2):  # second run: warm
            c = run(lib, mode, a.iters, a.grid)[0]
        per = (c - lat) / a.iters
        ideal = m * n * 16 * 2 / 8192
        print(f"{label:22s} {per:12.1f} {ideal:6.0f} {ideal / per:6.2f} {smem_bytes / per:10.1f}")
    if a.part == "pair":
        for mode, label in ((13, "remote mbarrier arrive round trip (2 hops)"),
                            (15, "remote arrive + pair MMA + multicast commit")):
            c = run(lib, mode, 64, 2)[0]
            print(f"{label:44s} {c / 64:8.1f} clk")
        return
    for mode, label, nw in ((16, "tcgen05.ld x32, 1 warp", 1), (8, "tcgen05.ld x32, 4 warps", 4),
                            (9, "tcgen05.ld x32, 8 warps", 8),
                            (10, "tcgen05.st x32, 4 warps", 4)):
        c = run(lib, mode, a.iters, a.grid)[0]
        nbytes = a.iters * 4 * 32 * 4 * 32 * nw  # groups x 4 instr x 32 cols x 4 B x 32 lanes x warps
        print(f"{label:28s} {c:9d} clk  {nbytes / c:7.1f} B/clk/SM")
    c = run(lib, 11, a.iters, a.grid)[0]
    print(f"{'MUFU ex2, 8 warps':28s} {c:9d} clk  {a.iters * 8 * 32 * 8 / c:7.2f} ex2/clk/SM")
    m, l, _, _ = run(lib, 14, a.iters, a.grid)
    print(f"TS chain beside 4 warps of tcgen05.ld: mma {(m - lat) / a.iters:.1f} clk/dispatch, "
          f"ld {a.iters * 4 * 32 * 4 * 32 * 4 / max(l, 1):.1f} B/clk/SM")
    if a.part == "all":
        for mode, label in ((13, "remote mbarrier arrive round trip (2 hops)"),
                            (15, "remote arrive + pair MMA + multicast commit")):
            c = run(lib, mode, 64, 2)[0]
            print(f"{label:44s} {c / 64:8.1f} clk")


if __name__ == "__main__":
    main()
