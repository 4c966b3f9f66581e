#!/usr/bin/env python
"""Probes of the CTA-pair (cta_group::2) mechanisms the pair backward needs, one GPU:
  * ba_selftest mode 7: B halves written by the peer's threads through DSMEM -> must equal A @ B^T;
  * ba_selftest mode 6: where an M = 128 cta_group::2 MMA puts its 64 rows per CTA in TMEM (raw dump, the
    rows that still hold the sentinel were not written).
    python tools/probe_pair.py > gpurun_out/probe_pair.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "burst-attention_b200"))
import torch  # noqa: E402

from burst_attn import native as nat  # noqa: E402


def selftest(mode, a, b):
    out = torch.zeros(256, 128, device=a.device, dtype=torch.float32)
    nat.check_selftest(nat.selftest_lib().ba_selftest(mode, a.data_ptr(), b.data_ptr(), out.data_ptr(), nat.dtype_code(a.dtype),
                                    nat.stream_ptr(a.device)), f"ba_selftest({mode})")
    torch.cuda.synchronize()
    return out


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    b = torch.randn(128, 128, device=dev).to(torch.bfloat16)
    # ---- mode 7
    a = torch.randn(256, 128, device=dev).to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    for mode in (4, 7):
        got = selftest(mode, a, b)
        err = (got - ref).abs().max().item()
        print(f"mode {mode} ({'TMA' if mode == 4 else 'DSMEM'}-delivered B halves): max abs err {err:.3e} "
              f"{'OK' if err < 1e-2 else 'MISMATCH'}")
    # ---- mode 6
    a = torch.randn(128, 128, device=dev).to(torch.bfloat16)
    ref = a.float() @ b.float().t()  # [128 rows, 128 cols]
    dump = selftest(6, a, b).view(2, 128, 128)
    for cta in range(2):
        owner = []
        for lane in range(128):
            row = dump[cta, lane]
            if torch.all(row == 12345.0):
                owner.append(None)
                continue
            d = (ref - row[None, :]).abs().amax(dim=1)
            r = int(d.argmin())
            owner.append(r if d[r] < 1e-2 else "?")
        # compress into runs
        runs, start = [], 0
        for i in range(1, 129):
            same = i < 128 and ((owner[i] is None and owner[start] is None) or
                                (isinstance(owner[i], int) and isinstance(owner[start], int)
                                 and owner[i] - i == owner[start] - start) or
                                (owner[i] == "?" and owner[start] == "?"))
            if not same:
                o = owner[start]
                what = "untouched" if o is None else ("unmatched" if o == "?" else f"rows {o}..{o + i - 1 - start}")
                runs.append(f"lanes {start}..{i - 1}: {what}")
                start = i
        print(f"mode 6, CTA {cta}: " + "; ".join(runs))


if __name__ == "__main__":
    main()
