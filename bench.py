#!/usr/bin/env python
"""bench.py -- attention TFLOPS/s (fwd+bwd, and fwd) at seq=262144, H=32, d=128, bf16,
bs=1 on N B200s (BASELINE.json metric), strong scaling: the global sequence is
fixed and sharded over the N ranks (contiguous shards, non-causal -- the C2/C3
configurations; N=1 is the local kernel with no ring).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference ...                     CPU arm: the oracle port of the
                                                           reference's path on the host cores

One "step" = one forward + one backward of burst_attn_func on synthetic
N(0,1) bf16 inputs already resident in HBM (`value`), and the same through the
public API from pinned HOST buffers with the H2D/D2H copies inside the timed
region (`e2e`).  FLOPs per benchmarks/benchmark.py:17-20 of the reference:
fwd 4*B*S^2*H*D, bwd 2.5x, fwd+bwd 3.5x.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "burst-attention_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

H, D, B = 32, 128, 1
METRIC = "attention fwd+bwd TFLOPS/s (bs=1, H=32, d=128, bf16; seq in config, default 262144), aggregate over GPUs"


def flops(S, mode, batch=None):
    f = 4.0 * (B if batch is None else batch) * S * S * H * D
    return {"fwd": f, "bwd": 2.5 * f, "fwd_bwd": 3.5 * f}[mode]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="MEASURED_PEAKS.json")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------- #
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- #
def cpu_port_step(S, threads):
    """One fwd+bwd of the reference's path restated by oracle/ (4 simulated ring
    rounds, fp32, torch CPU kernels on `threads` host threads).  Returns seconds."""
    from oracle import attention_oracle as orc
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    W, Hc = 4, 8
    q, k, v, do = (torch.randn(1, S, Hc, D, generator=g) for _ in range(4))
    sh = lambda t: [orc.shard(t, r, W, "contiguous") for r in range(W)]
    qs, ks, vs, dos = sh(q), sh(k), sh(v), sh(do)
    t0 = time.time()
    os_, lses = orc.ring_forward(qs, ks, vs, D ** -0.5, "none", torch.float32)
    orc.ring_backward(qs, ks, vs, os_, lses, dos, D ** -0.5, "none", torch.float32)
    dt = time.time() - t0
    return dt, 3.5 * 4.0 * S * S * Hc * D


def best_cpu_threads():
    """torch's CPU kernels do not scale to every hardware thread of a big host on these shapes (128
    threads were 10x slower than 8 here); probe a few thread counts on a small sample and keep the best."""
    n = os.cpu_count() or 1
    best, best_rate = n, 0.0
    for t in sorted({n, min(n, 64), min(n, 32), min(n, 16)}, reverse=True):
        cpu_port_step(512, t)
        dt, fl = cpu_port_step(2048, t)
        if fl / dt > best_rate:
            best, best_rate = t, fl / dt
    return best


def cpu_baseline(S=6144):
    threads = best_cpu_threads()
    dt, fl = cpu_port_step(S, threads)
    return {"value": fl / dt / 1e12, "unit": "TFLOPS/s", "cores": threads, "kind": "port",
            "sample": f"oracle port (torch CPU fp32), fwd+bwd, bs=1 S={S} H=8 d=128, 4 simulated ring rounds, "
                      f"{dt:.1f} s"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = best_cpu_threads()
    S = 4096
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_port_step(S, threads)
    steps = max(1, min(args.steps, 5))
    t = 0.0
    for _ in range(steps):
        dt, fl = cpu_port_step(S, threads)
        t += dt
    val = fl * steps / t / 1e12
    sample = (f"oracle port of the reference path (torch CPU fp32, {threads} threads): fwd+bwd bs=1 S={S} H=8 d=128, "
              f"4 simulated ring rounds per step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOPS/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "bounded CPU sample of the bench workload: " + sample},
        "cpu_baseline": {"value": val, "unit": "TFLOPS/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "TFLOPS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------- #
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seq", type=int, default=262144, help="global sequence length")
    ap.add_argument("--causal", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--configs", default="", help="comma list of extra runs in the same process group, e.g. "
                    "'262144,524288c,1048576' (c = causal zigzag); one JSON line each (multi-GPU sessions are "
                    "expensive to start)")
    ap.add_argument("--double-ring", type=int, default=0, metavar="L",
                    help="run over the hierarchical (double) ring with intra-node rings of L consecutive ranks "
                         "(reference benchmarks/benchmark.py --double_ring); default 0 = flat ring")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    W = max(3, args.warmup)
    K = max(1, args.steps)

    from burst_attn import burst_attn_func
    from burst_attn import chunk_ops, native
    native.check(native.lib().ba_device_check(), "ba_device_check")
    ops = chunk_ops.get_ops()

    args.double_group = [None, None]
    if args.double_ring and world > 1:
        L = args.double_ring
        assert world % L == 0 and 1 < L < world, "--double-ring L needs 1 < L < world and L | world"
        rows = [list(range(n * L, (n + 1) * L)) for n in range(world // L)]
        mk_groups = lambda ranks: dist.new_subgroups_by_enumeration(ranks, backend="nccl")[0]  # noqa: E731
        args.double_group = [mk_groups(rows), mk_groups([list(c) for c in zip(*rows)])]

    runs = [(args.seq, args.causal, B)]
    if args.configs:  # "<seq>[c][b<batch>]", e.g. 262144, 524288c, 65536b4 (the reference README's two sweeps)
        import re
        runs = []
        for c in args.configs.split(","):
            m = re.fullmatch(r"(\d+)(c?)(?:b(\d+))?", c.strip())
            assert m, f"bad config token {c!r}"
            runs.append((int(m.group(1)), bool(m.group(2)), int(m.group(3) or B)))
    for seq_i, causal_i, batch_i in runs:
        args.seq, args.causal, args.batch = seq_i, causal_i, batch_i
        _bench_one(args, world, rank, local, dev, W, K, ops, burst_attn_func)
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _bench_one(args, world, rank, local, dev, W, K, ops, burst_attn_func):
    S = args.seq
    Bn = getattr(args, "batch", B)
    S_loc = S // world
    layout = "zigzag" if args.causal else "contiguous"
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    mk = lambda: torch.randn(Bn, S_loc, H, D, device=dev, generator=gen, dtype=torch.float32).to(torch.bfloat16)
    q, k, v, do = mk(), mk(), mk(), mk()

    def step(qd, kd, vd, dod):
        qq, kk, vv = qd.detach().requires_grad_(), kd.detach().requires_grad_(), vd.detach().requires_grad_()
        o = burst_attn_func(qq, kk, vv, None, "cuda", args.causal, True, False, None, args.double_group)
        dq, dk, dv = torch.autograd.grad(o, (qq, kk, vv), dod)
        return o, dq, dk, dv

    def fwd_only(qd, kd, vd):
        with torch.no_grad():
            return burst_attn_func(qd, kd, vd, None, "cuda", args.causal, True, False, None, args.double_group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n

    # ---- warm-up (also builds the NCCL ring)
    for _ in range(W):
        step(q, k, v, do)
    torch.cuda.synchronize()

    # ---- timed: fwd+bwd, inputs resident in HBM; per-kernel events on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ops.enable_timing(True)
    launches0 = ops.launches
    t_wall0 = time.time()
    ms_step = timed(lambda: step(q, k, v, do), K)
    t_wall1 = time.time()
    launches = ops.launches - launches0
    torch.cuda.synchronize()
    kms = ops.kernel_ms()
    ops.enable_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_fwd = timed(lambda: fwd_only(q, k, v), max(1, min(K, 3)))

    causal_div = 2.0 if args.causal else 1.0
    fl_step = flops(S, "fwd_bwd", Bn) / causal_div
    value = fl_step / (ms_step * 1e-3) / 1e12
    fwd_tflops = flops(S, "fwd", Bn) / causal_div / (ms_fwd * 1e-3) / 1e12

    # ---- roofline of the dominant kernel (backward tile kernel: 2.5x the forward FLOPs)
    pk = peaks()
    roof = None
    if "bwd_chunk_kernel" in kms:
        n_l, tot_ms = kms["bwd_chunk_kernel"]
        # algorithmic FLOPs per launch = this rank's share of the step's backward FLOPs (5 GEMMs:
        # 10*Sq*Sk*H*D per round) / launches per step; non-causal: exactly 10*S_loc^2*H*D per ring round
        fl_launch = flops(S, "bwd", Bn) / causal_div / world / (n_l / K)
        ach = fl_launch / (tot_ms / n_l * 1e-3) / 1e12
        # DRAM traffic per launch from the committed ncu --set full capture (profiles/ncu_bwd_r01_final.txt:
        # dram__bytes_read 2.863 GB + write 1.771 GB) -- valid for the launch shape it was taken on
        # (Sq = Sk = 32768 per launch, H=32, non-causal: every multi-GPU ring round at S_local=32768)
        traffic = 4.634e9 if (S_loc == 32768 and Bn == 1 and not args.causal and n_l == K * world) else None
        roof = {"kernel": "bwd_chunk_kernel", "bound": "tensor", "achieved": ach, "peak": pk["sustained"],
                "unit": "TFLOP/s", "frac": ach / pk["sustained"], "traffic": traffic,
                "peak_source": pk["source"] + " bf16_tflops_sustained (of measured)",
                "launches": n_l, "avg_launch_ms": tot_ms / n_l}
        if "fwd_chunk_kernel" in kms:
            n_f, tot_f = kms["fwd_chunk_kernel"]
            fl_f = flops(S, "fwd", Bn) / causal_div / world / (n_f / K)
            roof["fwd_kernel"] = {"achieved": fl_f / (tot_f / n_f * 1e-3) / 1e12, "launches": n_f,
                                  "avg_launch_ms": tot_f / n_f,
                                  "frac": fl_f / (tot_f / n_f * 1e-3) / 1e12 / pk["sustained"]}

    # ---- how much of the step is NOT inside one of our kernels on the compute stream: torch memsets /
    # allocations, launch gaps and any ring-communication time the kernels did not hide (upper bound
    # on exposed comm; target < 5 %)
    overlap = None
    if kms:
        k_ms = sum(t for _, t in kms.values()) / K
        overlap = {"kernel_ms_per_step": k_ms, "non_kernel_ms_per_step": ms_step - k_ms,
                   "non_kernel_frac": (ms_step - k_ms) / ms_step,
                   "per_kernel_ms_per_step": {n: t / K for n, (c, t) in kms.items()}}

    # ---- e2e: same step through the public API from pinned host buffers
    e2e = None
    if not args.no_e2e:
        hq, hk, hv, hdo = (t.cpu().pin_memory() for t in (q, k, v, do))
        ho = [torch.empty_like(hq).pin_memory() for _ in range(4)]

        def e2e_step():
            dq_, dk_, dv_, ddo_ = (h.to(dev, non_blocking=True) for h in (hq, hk, hv, hdo))
            outs = step(dq_, dk_, dv_, ddo_)
            for h, t in zip(ho, outs):
                h.copy_(t, non_blocking=True)
        e2e_step()
        ms_e2e = timed(e2e_step, K)
        nbytes = hq.numel() * hq.element_size()
        e2e = {"value": fl_step / (ms_e2e * 1e-3) / 1e12, "unit": "TFLOPS/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": 4 * nbytes * world, "d2h_bytes_per_step": 4 * nbytes * world}

    tot_launch = torch.tensor([launches], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(tot_launch)

    if rank == 0:
        cpu = None if (args.no_cpu or world > 1) else cpu_baseline()  # reported on rank 0 at N=1 only
        # the reference's README publishes 191 TFLOPS/s/GPU fwd+bwd at S=262144 on 8 GPUs (8xA100)
        vs = value / (191.0 * 8) if (world == 8 and S == 262144 and not args.causal) else None
        line = {
            "metric": METRIC, "value": value, "unit": "TFLOPS/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": vs,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"burst_attn_func fwd+bwd, bs={Bn} S={S} (S_local={S_loc}) H=32 d=128 bf16 "
                                   f"{'causal zigzag' if args.causal else 'non-causal contiguous'} shards, "
                                   f"{'local kernel, no ring' if world == 1 else f'{world}-rank ring over NCCL'}"
                                   f"{f' (double ring, intra {args.double_ring})' if args.double_ring and world > 1 else ''}",
                       "global_batch": Bn, "seq_len": S, "parallelism": f"sp{world}",
                       "l2": "inputs (>= 256 MiB per tensor per rank) exceed the 126 MB L2; no flush needed"},
            "value_per_gpu": value / world, "fwd_tflops": fwd_tflops, "fwd_ms": ms_fwd,
            "gpu_launches": int(tot_launch.item()), "clocks": clocks, "e2e": e2e, "roofline": roof, "overlap": overlap,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
