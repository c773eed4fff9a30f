#!/usr/bin/env python
"""bench.py -- separated-audio-seconds per second of the Conv-TasNet path (forward + SI-SDR/PIT loss).

    python bench.py --gpus N --steps K --warmup W            # our sm_100a path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores (oracle port)

Workload (BASELINE.json configs[1], "cfg2"): Conv-TasNet N=512 L=16 B=128 H=512 Sc=128 P=3 X=8 R=3, gLN, 2 speakers,
batch 32 x 4 s @ 8 kHz per GPU (weak scaling: every rank gets its own batch of 32; no data-path collective).
A "step" = one pass of the hot path (model forward + PIT(NegSISDR) loss) over one batch of synthetic mixtures.
Prints ONE JSON line on rank 0.  value = device-resident throughput; e2e = through the public module API with pinned
host inputs, H2D copies and the D2H loss/permutation read inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dnn-based_source_separation_b200"))

SR = 8000
PAPER = dict(n_basis=512, kernel_size=16, sep_hidden_channels=512, sep_bottleneck_channels=128, sep_skip_channels=128,
             sep_kernel_size=3, sep_num_blocks=3, sep_num_layers=8)
METRIC = "audio-sec/s Conv-TasNet 2spk 4s@8kHz fwd+SI-SDR-PIT"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="mixtures per GPU per step")
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--sample-rate", type=int, default=8000, help="Hz; cfg5 of BASELINE.json is 8 s @ 16 kHz")
    ap.add_argument("--n-sources", type=int, default=2)
    ap.add_argument("--math", default=None, choices=[None, "fp32", "tf32x3", "tf32", "f16x3"])
    ap.add_argument("--cpu-batch", type=int, default=4, help="mixtures per CPU-baseline step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train", action="store_true",
                    help="time the TRAINING step instead (fwd + PIT + backward + gradient all-reduce + clip + Adam); prints its own line")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.th = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(args, steps, warmup):
    """The reference algorithm (oracle port, plain PyTorch CPU ops = what the reference executes) on all host cores,
    on a bounded sample of the workload: cpu-batch mixtures of the same 4 s @ 8 kHz shape per step."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import convtasnet_oracle as O
    cores = os.cpu_count() or 1
    cfg = O.OracleConfig(**PAPER, causal=False, n_sources=args.n_sources)
    sd = O.synth_state_dict(cfg, seed=111)
    T = int(args.seconds * args.sample_rate)
    mixture, sources = O.synth_batch(args.cpu_batch, args.n_sources, T, seed=111)
    # "all the host threads it can use": torch's intra-op pool saturates well below 128 threads on these tensor sizes and
    # gets SLOWER beyond that, so sweep a few team sizes on one sample and keep the fastest (reported as `cores`).
    best_thr, best_t = cores, float("inf")
    sweep = sorted({t for t in (8, 16, 32, 64, cores) if t <= cores})
    with torch.no_grad():
        for thr in sweep:
            torch.set_num_threads(thr)
            O.conv_tasnet_fwd(mixture[:1], sd, cfg)
            t0 = time.perf_counter()
            O.conv_tasnet_fwd(mixture[:1], sd, cfg)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best_thr, best_t = thr, dt
    torch.set_num_threads(best_thr)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out, _ = O.conv_tasnet_fwd(mixture, sd, cfg)
            loss, perm = O.pit_neg_sisdr(out, sources)
            float(loss)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    value = args.cpu_batch * args.seconds * len(times) / total
    return dict(value=value, unit="audio-sec/s", cores=best_thr, threads=torch.get_num_threads(), kind="port",
                sample=f"{args.cpu_batch} x {args.seconds:g} s @ {args.sample_rate} Hz per step, {len(times)} steps (+{warmup} warm-up), "
                       f"oracle/convtasnet_oracle.py fwd+PIT under no_grad, {best_thr} torch threads (fastest of {sweep}) on "
                       f"{cores} logical cores",
                ms_per_step=1e3 * total / len(times))


def stage_model(args, B, frames, T):
    """Algorithmic (bytes, flops) per LAUNCH of each stage."""
    N, Bc, H, Sc, S = PAPER["n_basis"], PAPER["sep_bottleneck_channels"], PAPER["sep_hidden_channels"], PAPER["sep_skip_channels"], args.n_sources
    L = PAPER["kernel_size"]
    f = frames * B * 4.0
    Mt = Bc + Sc
    return {
        "enc": (B * T * 4.0 + N * f, 2.0 * N * L * frames * B, "hbm"),
        "head": ((N + Bc) * f, 2.0 * N * Bc * frames * B, "tensor"),
        "pw1": ((Bc + H) * f, 2.0 * Bc * H * frames * B, "tensor"),
        "dw": (2.0 * H * f, 2.0 * 3 * H * frames * B, "hbm"),
        "pw2": ((H + Mt) * f, 2.0 * H * Mt * frames * B, "tensor"),
        "fin": (3.0 * Mt * f, 2.0 * Mt * frames * B, "hbm"),
        "mask": ((Sc + N + S * N) * f, 2.0 * Sc * S * N * frames * B, "tensor"),
        "dec": (S * N * f + S * B * T * 4.0, 2.0 * S * N * L * frames * B, "hbm"),
        "loss": (2 * 2.0 * S * B * T * 4.0 / 3.0, 0.0, "hbm"),   # 3 launches share two passes over est+tgt
        "prep": (0.0, 0.0, "hbm"),
    }


def main():
    args = parse()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        leg = cpu_reference_leg(args, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": leg["value"], "unit": "audio-sec/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"Conv-TasNet {args.n_sources}spk N512 L16 B128 H512 Sc128 P3 X8 R3 gLN, {args.seconds:g}s@8kHz, "
                                       f"fwd+SI-SDR-PIT; CPU step = {args.cpu_batch} mixtures (bounded sample of the batch-{args.batch} workload)"},
                "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": leg["value"], "unit": "audio-sec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    from ctn_b200 import _native as N
    from ctn_b200 import dist as D
    from ctn_b200.models.conv_tasnet import ConvTasNet
    from ctn_b200.criterion.sdr import NegSISDR
    from ctn_b200.criterion.pit import PIT1d

    rank, local_rank, world = D.init()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    S, B, T = args.n_sources, args.batch, int(args.seconds * args.sample_rate)

    torch.manual_seed(111)  # reference default seed (train.sh:59); default init = the reference's default init
    model = ConvTasNet(PAPER["n_basis"], PAPER["kernel_size"], enc_basis="trainable", dec_basis="trainable", enc_nonlinear=None,
                       sep_hidden_channels=PAPER["sep_hidden_channels"], sep_bottleneck_channels=PAPER["sep_bottleneck_channels"],
                       sep_skip_channels=PAPER["sep_skip_channels"], sep_kernel_size=3, sep_num_blocks=3, sep_num_layers=8,
                       causal=False, n_sources=S).to(dev).eval()
    model.math = args.math
    math_name = args.math or ("f16x3" if N.ctn_has_tcgen05() else "fp32")
    crit = PIT1d(NegSISDR(), S)
    g = torch.Generator().manual_seed(111 + rank)
    sources_h = (0.1 * torch.randn(B, S, T, generator=g)).pin_memory()
    mixture_h = sources_h.sum(dim=1, keepdim=True).pin_memory()
    mixture_d, sources_d = mixture_h.to(dev), sources_h.to(dev)
    frames = N.frames_of(T, PAPER["kernel_size"], PAPER["kernel_size"] // 2)[0]

    def step_resident():
        out = model(mixture_d)
        return crit(out, sources_d)

    loss_pin = torch.empty(1).pin_memory()
    perm_pin = torch.empty(B, S, dtype=torch.int64).pin_memory()

    def step_e2e():
        x = mixture_h.to(dev, non_blocking=True)
        t = sources_h.to(dev, non_blocking=True)
        out = model(x)
        loss, perm = crit(out, t)
        loss_pin.copy_(loss.reshape(1), non_blocking=True)
        perm_pin.copy_(perm, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller reads loss / perm every step (driver.py:157 loss.item())
        return float(loss_pin[0])

    if args.train:
        # ---- training step (driver.py:146-157): fwd_train + PIT + native backward + ONE gradient all-reduce + clip + Adam ----
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
        nelem = 0

        def step_train():
            nonlocal nelem
            opt.zero_grad(set_to_none=True)
            loss, _ = crit(model(mixture_d), sources_d)
            loss.backward()
            nelem = D.allreduce_gradients(model)
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
            opt.step()
            return loss

        for _ in range(max(args.warmup, 2)):
            loss = step_train()
        launches = model.last_launches + model.last_bwd_launches + 3
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record()
            for _ in range(args.steps):
                loss = step_train()
            e1.record()
            torch.cuda.synchronize()
        D.barrier()
        ms = D.max_over_ranks(e0.elapsed_time(e1), dev)
        peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
        if rank == 0:
            print(json.dumps({
                "mode": "train", "metric": "audio-sec/s Conv-TasNet %dspk %gs@8kHz TRAIN step (fwd+SI-SDR-PIT+bwd+allreduce+clip+Adam)" % (S, args.seconds),
                "value": world * B * args.seconds * args.steps / (ms * 1e-3), "unit": "audio-sec/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 2), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "dtype": math_name, "data": "synthetic", "config": {"workload": f"cfg2/cfg3 shape, batch {B} per GPU", "global_batch": world * B,
                "optimizer": "torch.optim.Adam(fused) + clip_grad_norm_ (torch; not part of the native path)"},
                "gpu_launches": launches * args.steps, "allreduce_elems": nelem, "peak_mem_gb": peak_gb, "clocks": clk.summary(),
                "last_loss": float(loss)}), flush=True)
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    launches_per_step = 0
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            loss, perm = step_resident()
        launches_per_step = model.last_launches + N.ctn_last_launch_count()
        torch.cuda.synchronize()

        # ---- timed: device-resident ---------------------------------------------------------------------------
        N.ctn_profile_enable(1)
        N.profile_read()
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            e0.record()
            for _ in range(args.steps):
                loss, perm = step_resident()
            e1.record()
            torch.cuda.synchronize()
        D.barrier()
        ms_local = e0.elapsed_time(e1)
        prof = N.profile_read()
        N.ctn_profile_enable(0)
        ms = D.max_over_ranks(ms_local, dev)

        # ---- timed: end to end --------------------------------------------------------------------------------
        for _ in range(2):
            step_e2e()
        D.barrier()
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(args.steps):
            last_loss = step_e2e()
        e3.record()
        torch.cuda.synchronize()
        D.barrier()
        ms_e2e = D.max_over_ranks(e2.elapsed_time(e3), dev)

    audio_per_step = world * B * args.seconds
    value = audio_per_step * args.steps / (ms * 1e-3)
    e2e_value = audio_per_step * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        return

    pk = peaks()
    model_bf = stage_model(args, B, frames, T)
    # TF32 dense = 1/2 bf16 on tcgen05; fp16 operands (f16x3) run at the bf16 rate.  Sustained figure (kernel timed inside a long step)
    tf32_peak = pk["bf16_sustained"] / (1.0 if math_name == "f16x3" else 2.0)
    stages = {}
    for name, (t_ms, n) in prof.items():
        if n == 0:
            continue
        by, fl, bound = model_bf[name]
        per_launch_ms = t_ms / n
        groups = max(1, {"pw1": 24, "dw": 24, "pw2": 24, "fin": 24}.get(name, 1) * args.steps)
        # stage records are per kernel group (one per block per step); bytes/flops above are per group
        per_group_ms = t_ms / groups if name != "prep" else t_ms / args.steps
        ent = {"ms_per_step": t_ms / args.steps, "launches_per_step": n / args.steps, "share": t_ms / (ms_local + 1e-9)}
        if by > 0:
            ent["GBps"] = by / (per_group_ms * 1e-3) / 1e9
            ent["hbm_frac"] = ent["GBps"] / pk["hbm"]
        if fl > 0 and bound == "tensor":
            ent["TFLOPs"] = fl / (per_group_ms * 1e-3) / 1e12
            ent["tf32_frac"] = ent["TFLOPs"] / tf32_peak
        ent["bound"] = bound
        ent["avg_launch_ms"] = per_launch_ms
        stages[name] = ent
    dom = max((k for k in stages if k != "prep"), key=lambda k: stages[k]["ms_per_step"])
    d = stages[dom]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(math_name, {}).get(dom)
    if d["bound"] == "tensor":
        roof = {"kernel": dom, "bound": "tensor", "achieved": d["TFLOPs"], "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": d["TFLOPs"] / tf32_peak, "traffic": traffic,
                "peak_note": (f"fp16 dense = bf16_tflops_sustained of {pk['source']}" if math_name == "f16x3" else
                              f"TF32 dense = bf16_tflops_sustained/2 of {pk['source']}") +
                             "; algorithmic 2*M*N*K flops (the 3-pass hi/lo split issues 3x that on the tensor pipe)"}
        # the same kernel against the other roofline (algorithmic bytes / duration): with the fp16 pieces the two ideal times
        # are within 15 % of each other, so both fractions are reported
        if d.get("GBps") is not None:
            roof["hbm_view"] = {"achieved": d.get("GBps"), "peak": pk["hbm"], "unit": "GB/s", "frac": d.get("hbm_frac")}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": d["GBps"], "peak": pk["hbm"], "unit": "GB/s",
                "frac": d["GBps"] / pk["hbm"], "traffic": traffic, "peak_note": f"hbm_gbs of {pk['source']}"}

    line = {
        "metric": METRIC, "value": value, "unit": "audio-sec/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32 (CUDA-core FFMA)", "tf32x3": "f32 via 3xTF32 split on tcgen05, fp32 accumulate", "tf32": "tf32 (single pass), fp32 accumulate", "f16x3": "f32 via 3xFP16 split on tcgen05 (kind::f16), fp32 accumulate"}[math_name],
        "data": "synthetic",
        "config": {"workload": f"cfg2: Conv-TasNet {S}spk N512 L16 B128 H512 Sc128 P3 X8 R3 gLN sigmoid, batch {B} x {args.seconds:g}s@{args.sample_rate // 1000}kHz per GPU, "
                               f"fwd + PIT(NegSISDR)", "global_batch": world * B, "math": math_name,
                   "l2": "per-step activation traffic (>2 GB) exceeds the 126 MB L2 many times over; no explicit flush",
                   "parallelism": f"batch shards x{world}, no data-path collective"},
        "e2e": {"value": e2e_value, "unit": "audio-sec/s", "h2d_bytes_per_step": B * T * 4 * (1 + S), "d2h_bytes_per_step": 4 + B * S * 8,
                "ms_per_step": ms_e2e / args.steps, "api": "ConvTasNet.forward + PIT1d(NegSISDR).forward on pinned host tensors"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roof, "stages": stages, "clocks": clk.summary(), "last_loss": last_loss,
    }
    if world == 1 and not args.no_cpu_baseline:
        leg = cpu_reference_leg(args, steps=3, warmup=1)
        line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
