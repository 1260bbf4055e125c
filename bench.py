"""Benchmark of the EaBNet inference hot path (BASELINE.json: audio-seconds enhanced per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch 64] [--seconds 6]

One step = one wave -> wave pass (STFT+compression, network, filter-and-sum, iSTFT) over one batch of synthetic
9-mic utterances: BASELINE.json configs[1] (64 x 6 s on one B200).  N > 1 (torchrun, one rank per GPU) shards
independent utterance batches over the ranks with no collective on the data path ("weak" scaling: every rank
enhances its own 64 x 6 s batch per step); timing is CUDA events on the launching stream bracketed by barriers,
max over ranks.

`value`  : throughput with the input waveforms already resident in HBM (eab_enhance, device pointers).
`e2e`    : the same work with HOST (pinned) buffers through eab_enhance_host_batches (`steps` batches per call): H2D of
           every batch's waveforms and D2H of its enhanced audio inside the timed region, overlapped with compute.
`roofline`: the kernel family with the largest share of the step, timed live per launch with CUDA events, against SURVEY.md
           section 8(d)'s ALGORITHMIC bytes (every fused layer reads its fp32 inputs once and writes its output once: 2-D convs
           674 888 B + TCM 73 728 B + head 301 392 B per frame = 40.4 GB per 64 x 6 s step); what the design actually moves
           is reported next to it (`moved_bytes`, `traffic_ratio`), as are all families and the whole step.
`gpu_eager_baseline`: the oracle port run on the SAME GPU through torch eager (cuDNN / cuFFT / cuBLAS, TF32 off) - BASELINE.md
           section 3.4's "real bar to beat"; `single_utterance` = BASELINE configs[0] shape (B = 1, 4 s) on the GPU;
           `config4` = BASELINE configs[3] (2048 x 6 s through shard_range + enhance_host_batches, fixed total size).
`cpu_baseline` / `--impl reference`: the CPU oracle port of the reference path (oracle/eabnet_oracle.py, torch
           fp32 ops, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio_seconds_enhanced_per_second"
DTYPE = "f32 io/accum; fp16x3 (hi/lo split) enc/TCM/LSTM/head/STFT, fp16x1 decoder"
WORKLOAD = ("EaBNet default (M=9, causal, U2, LSTM head, IN) batched wave->wave enhancement, "
            "64 x 6 s utterances per GPU (BASELINE configs[1])")
# algorithmic bytes per frame of the layer-fused path (fp32, inputs once + output once per layer)
ALGO_BYTES_PER_FRAME = {"conv2d": 674888.0, "tcm": 73728.0, "head": 301392.0}
# kernel families: profiler categories -> SURVEY rows (a3-a7/a11 2-D conv stack, a10 TCM stack, a12-a13 head, a1, a14)
FAMILY_OF = {"conv_raw": "conv2d", "conv_tma": "conv2d", "stage": "conv2d", "conv_umma": "conv2d", "conv_generic": "conv2d",
             "combine": "conv2d", "tcm_chain": "tcm", "lstm_umma": "head", "lstm": "head", "head_fused": "head", "beam": "head",
             "stft": "stft", "stft_stage": "stft", "istft": "istft"}
UNIT = "audio-s/s"
SR = 16000


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled while the timed region runs: in-process NVML (cheap) every 100 ms,
    falling back to an `nvidia-smi -lms` subprocess when the NVML binding is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc, self._stop_evt = index, [], None, threading.Event()
        self.source = "nvml"

    def _run_nvml(self) -> bool:
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            return False
        while not self._stop_evt.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx), "0"] + ["Active" if mask & self.BITS[n] else "Not Active" for n in self.NAMES])
            except Exception:
                pass
            self._stop_evt.wait(0.1)
        return True

    def run(self):
        if self._run_nvml():
            return
        self.source = "nvidia-smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self) -> dict:
        self._stop_evt.set()
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        for r in list(self.rows):
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(self.NAMES, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": self.source}


def cpu_reference_throughput(batch: int, seconds: float, steps: int, warmup: int):
    """The reference path on the host cores: oracle port (torch fp32, all threads), wave -> wave."""
    import torch
    from oracle import eabnet_oracle as O          # the one place bench.py executes oracle/: the CPU baseline
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.make_cfg()
    sd = O.make_weights(cfg, 0, "B")
    L = int(seconds * SR)
    wave, _ = O.make_wave(batch, 9, L, seed=1234)
    for _ in range(warmup):
        O.enhance(sd, wave, cfg)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.enhance(sd, wave, cfg)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return batch * seconds / med, med, cores, "%d x %.0f s 9-mic utterances per step, median of %d steps" % (batch, seconds, steps)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    val, med, cores, sample = cpu_reference_throughput(args.ref_batch, args.seconds, steps, max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": args.batch, "seconds": args.seconds, "sample_batch": args.ref_batch,
                       "caps": "CPU arm: at most 5 timed steps and 1 warm-up of an %d x %.0f s sample (about 12 s each on 16 cores), "
                               "whatever --steps / --warmup ask for; value = median step" % (args.ref_batch, args.seconds)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def stream_latency(args, dev):
    """BASELINE configs[2]: `streams` concurrent causal streams, one 10 ms hop (160 samples x 9 mics) per step with
    carried conv / TCM / LSTM state (norm_type='BN': InstanceNorm spans the utterance), through eab_stream_step replayed
    from a CUDA graph.  Host-to-host: pinned hop in, enhanced hop back, wall clock per step."""
    import torch
    from eabnet_b200 import EaBNet
    from eabnet_b200.model import EaBNetStream
    S, N = args.streams, args.stream_steps
    torch.manual_seed(7)
    net = EaBNet(norm_type="BN").eval().to(dev)
    ses = EaBNetStream(net, S, dev, graph=True)
    hop_host = (0.1 * torch.randn(S, 9, 160)).pin_memory()
    out_host = torch.empty(S, 160).pin_memory()
    hop = torch.empty(S, 9, 160, device=dev)
    out = torch.empty(S, 160, device=dev)
    for _ in range(20):
        ses.step(hop, out)
    torch.cuda.synchronize(dev)
    wall, devms = [], []
    for _ in range(N):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        hop.copy_(hop_host, non_blocking=True)
        e0.record()
        ses.step(hop, out)
        e1.record()
        out_host.copy_(out, non_blocking=True)
        torch.cuda.synchronize(dev)
        wall.append((time.perf_counter() - t0) * 1e3)
        devms.append(e0.elapsed_time(e1))
    wall.sort()
    devms.sort()
    assert torch.isfinite(out_host).all()
    return {"workload": "causal streaming, %d concurrent 9-mic streams, one 10 ms hop per step, carried state, BN "
                        "(BASELINE configs[2])" % S,
            "streams": S, "steps": N, "hop_ms": 10.0, "p50_ms": wall[N // 2], "p99_ms": wall[min(N - 1, int(N * 0.99))],
            "device_p50_ms": devms[N // 2], "launches_per_step": net.last_launch_count(), "cuda_graph": True,
            "timing": "wall clock per step: pinned H2D of the hop + step + D2H of the enhanced hop + sync",
            "realtime_factor": S * 10.0 / wall[N // 2]}


def postnet_throughput(args, dev, wave):
    """what enhance.py actually runs - EaBNetWithPostNet (EaBNet + GaGNet post-filter) wave -> wave
    on the same synthetic batch, through eab_enhance_postnet; device-resident inputs, CUDA events."""
    import torch
    from eabnet_b200 import make_eabnet_with_postnet
    from eabnet_b200.postnet import default_postnet_args
    torch.manual_seed(4321)
    w = make_eabnet_with_postnet(default_postnet_args()).eval().to(dev)
    steps = max(3, min(args.steps, 10))
    with torch.no_grad():
        g = w.graphed_enhance(wave) if not args.no_graph else None
        run_step = g.step if g is not None else (lambda: w.enhance(wave))
        for _ in range(3):
            y = run_step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            y = run_step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        launches = g.launches if g is not None else w.eabnet.last_launch_count()
        w.postnet.profile(True)
        w.enhance(wave)
        prof = w.postnet.profile_summary()
        w.postnet.profile(False)
    assert torch.isfinite(y).all()
    B, L = wave.shape[0], wave.shape[2]
    return {"workload": "EaBNetWithPostNet (EaBNet + GaGNet post-filter, enhance.py:21,49-62) wave->wave, %d x %.0f s per GPU" % (B, L / SR),
            "value": B * (L / SR) / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "launches_per_step": launches,
            "kernels": prof}


def gpu_eager_baseline(args, dev):
    """the reference's own op sequence in PyTorch eager on this GPU (cuDNN convs, cuFFT, cuBLAS, the
    fused cuDNN LSTM), TF32 off so that it computes what the CPU reference computes.  The oracle port stands in for the
    reference module (same torch primitives; the module itself does not travel to the GPU box).  Bounded sample."""
    import torch
    from oracle import eabnet_oracle as O          # baseline leg (checker / timed baseline only)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        cfg = O.make_cfg()
        sd = {k: v.to(dev) for k, v in O.make_weights(cfg, 0, "B").items()}
        Bs, L = args.ref_batch, int(args.seconds * SR)
        wave = O.make_wave(Bs, 9, L, seed=1234)[0].to(dev)
        with torch.no_grad():
            for _ in range(2):
                y = O.enhance(sd, wave, cfg)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                y = O.enhance(sd, wave, cfg)
            e1.record()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 5
        assert torch.isfinite(y).all()
        return {"value": Bs * args.seconds / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "kind": "port",
                "what": "oracle port in torch eager on cuda (cuDNN / cuFFT / cuBLAS, allow_tf32 = False), wave -> wave, device-resident",
                "sample": "%d x %.0f s 9-mic utterances per step, mean of 5 steps after 2 warm-ups" % (Bs, args.seconds)}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def single_utterance(dev, net):
    """BASELINE configs[0] on the GPU: one 4 s 9-mic utterance, the shape enhance.py feeds (B = 1, T = 401)."""
    import torch
    from eabnet_b200 import stft_compress
    L = 64000
    g = torch.Generator().manual_seed(5)
    wave_host = (0.1 * torch.randn(1, 9, L, generator=g)).pin_memory()
    out_host = torch.empty(1, 160 * (L // 160)).pin_memory()
    wave = wave_host.to(dev)
    with torch.no_grad():
        spec = stft_compress(wave)
        fwd, host = [], []
        for i in range(60):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = net(spec)
            e1.record()
            torch.cuda.synchronize(dev)
            if i >= 10:
                fwd.append(e0.elapsed_time(e1))
        launches = net.last_launch_count()
        for i in range(60):
            t0 = time.perf_counter()
            net.enhance_host(wave_host, out_host, dev)
            torch.cuda.synchronize(dev)
            if i >= 10:
                host.append((time.perf_counter() - t0) * 1e3)
    fwd.sort()
    host.sort()
    assert torch.isfinite(y).all() and torch.isfinite(out_host).all()
    return {"workload": "one 4 s 9-mic utterance (BASELINE configs[0] shape, B = 1, T = 401)",
            "forward_p50_ms": fwd[len(fwd) // 2], "forward_launches": launches,
            "enhance_host_p50_ms": host[len(host) // 2], "enhance_host_p99_ms": host[min(len(host) - 1, int(len(host) * 0.99))],
            "audio_s_per_s_host": 4.0 / (host[len(host) // 2] * 1e-3),
            "timing": "forward: CUDA events around net(noisy_stft), p50 of 50; enhance_host: wall clock around the host-buffer call "
                      "(H2D + STFT + net + iSTFT + D2H + sync), p50 of 50"}


def config4(args, dev, net, world, rank, ins2, outs2):
    """BASELINE configs[3]: 2048 x 6 s utterances (fixed total), batches of 64 dealt to the ranks by shard_range, each rank
    pushing its batches through enhance_host_batches (host buffers, copies overlapped); wall clock, max over ranks."""
    import torch
    from eabnet_b200.shard import max_over_ranks, shard_range
    total_utt = 2048
    nb_total = (total_utt + args.batch - 1) // args.batch
    b0, b1 = shard_range(nb_total, rank, world)
    n = b1 - b0
    ins = [ins2[i % 2] for i in range(n)]
    outs = [outs2[i % 2] for i in range(n)]
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    if n:
        net.enhance_host_batches(ins, outs, dev)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    (dt,) = max_over_ranks([dt], dev)
    return {"workload": "2048 x %.0f s 9-mic utterances in batches of %d, dealt to %d rank(s) with shard_range, host buffers "
                        "(BASELINE configs[3]); synthetic batches re-used from two pinned buffers" % (args.seconds, args.batch, world),
            "utterances": nb_total * args.batch, "batches_this_rank": n, "seconds_total": dt, "scaling": "strong",
            "value": nb_total * args.batch * args.seconds / dt, "unit": UNIT}


def config5_training(args, dev, world, rank, local):
    """BASELINE configs[4]: the training step of train_distributed.py:214-230 (batch 16 x 4 s per GPU, data-parallel, NCCL gradient
    allreduce).  The backward pass of the product is NOT built beyond its first slice, so this object reports (a) the stated
    baseline: the oracle port of the whole step in torch eager on this GPU (cuDNN / cuBLAS autograd), wrapped in
    DistributedDataParallel when launched under torchrun, with the 35.16 MB gradient allreduce timed on its own; (b) the slice of
    the step the product's hand-written kernels cover today (head tail + loss, forward and backward) at the same shapes."""
    import torch
    import torch.distributed as dist
    from oracle import eabnet_oracle as O          # baseline leg only
    from oracle import train_oracle as TO
    Bt, Lt = args.train_batch, int(4.0 * SR)
    torch.manual_seed(7 + rank)
    model = TO.TrainableEaBNetWithPostNet().to(dev)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=5e-4)
    wave, clean = O.make_wave(Bt, 9, Lt, seed=500 + rank)
    wave, clean = wave.to(dev), clean.to(dev)

    def step():
        opt.zero_grad()
        noisy = O.stft_compress(wave)                                            # prepare_data, test.py:20-47
        target = O.stft_compress(clean.unsqueeze(1))[..., 0, :].permute(0, 3, 1, 2).contiguous()
        out = net(noisy)
        l = TO.loss_fn(out, target, [noisy.shape[1]] * Bt)
        l["final"].backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        return l["final"].detach(), noisy, target, out

    loss, noisy, target, out = step()                                            # warm-up (cuDNN plans, allocator)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nst = 3
    e0.record()
    for _ in range(nst):
        loss, noisy, target, out = step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms_step = e0.elapsed_time(e1) / nst
    nparam = sum(p.numel() for p in model.parameters() if p.requires_grad)
    res = {"workload": "train_distributed.py:214-230 training step, EaBNetWithPostNet, batch %d x 4 s per GPU, %d rank(s)%s (BASELINE configs[4])"
                       % (Bt, world, ", DistributedDataParallel over NCCL" if world > 1 else ""),
           "baseline": {"kind": "port", "what": "oracle port of the training step in torch eager on cuda (autograd, Adam, clip_grad_norm_)",
                        "ms_per_step": ms_step, "audio_s_per_s": world * Bt * 4.0 / (ms_step * 1e-3), "steps": nst,
                        "loss": float(loss), "grad_bytes": nparam * 4}}
    if world > 1:
        flat = torch.zeros(nparam, dtype=torch.float32, device=dev)
        for _ in range(2):
            dist.all_reduce(flat)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(10):
            dist.all_reduce(flat)
        e1.record()
        torch.cuda.synchronize(dev)
        ar = e0.elapsed_time(e1) / 10
        res["baseline"]["allreduce_ms"] = ar
        res["baseline"]["allreduce_busbw_GBps"] = 2.0 * (world - 1) / world * nparam * 4 / (ar * 1e-3) / 1e9
    if rank == 0:
        # the product's slice at the same shapes: h2 = a stand-in for the second LSTM's output, the real spectrum and target
        from eabnet_b200.train import com_mag_mse_loss, head_filter_sum
        sdm = dict(zip(model.names, model.values))
        Wn = ["eabnet.bf_map.w_dnn.0.weight", "eabnet.bf_map.w_dnn.0.bias", "eabnet.bf_map.w_dnn.2.weight", "eabnet.bf_map.w_dnn.2.bias"]
        Ws = [sdm[k].detach().clone().requires_grad_(True) for k in Wn]
        T = noisy.shape[1]
        h2 = torch.tanh(torch.randn(Bt, T, 161, 64, device=dev)).requires_grad_(True)
        spec = noisy.detach()

        def slice_step():
            o = head_filter_sum(h2, spec, *Ws)
            lo = com_mag_mse_loss(o, target)
            return torch.autograd.grad(lo, [h2] + Ws)
        for _ in range(2):
            slice_step()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(5):
            slice_step()
        e1.record()
        torch.cuda.synchronize(dev)
        res["product_slice"] = {"what": "hand-written kernels: w_dnn + filter-and-sum forward / backward (d h2, dW1, db1, dW2, db2) and "
                                        "com_mag_mse_loss forward / backward (csrc/head_bwd.cu)", "ms": e0.elapsed_time(e1) / 5,
                                "rows": Bt * T * 161, "note": "everything upstream of h2 (LSTM, decoder, TCMs, encoder, STFT) has no backward kernel yet"}
    del model, net, opt
    torch.cuda.empty_cache()
    return res


def build_roofline(prof, ms_step, frames, peaks):
    """Per-family and whole-step roofline fractions on SURVEY 8(d) algorithmic bytes; `prof` = one profiled step."""
    fam = {}
    for k in prof:
        f = FAMILY_OF.get(k["kernel"], "other")
        d = fam.setdefault(f, {"ms": 0.0, "flops": 0.0, "launcher_bytes": 0.0, "moved_bytes": 0.0, "launches": 0, "kernels": {}})
        d["ms"] += k["ms"]; d["flops"] += k["flops"]; d["launcher_bytes"] += k["bytes"]
        d["moved_bytes"] += k.get("moved_bytes", k["bytes"]); d["launches"] += k["launches"]
        d["kernels"][k["kernel"]] = round(k["ms"], 4)
    hbm, tens = peaks["hbm_gbs"], peaks["bf16_tflops_sustained"]
    out = {}
    for f, d in fam.items():
        algo = ALGO_BYTES_PER_FRAME[f] * frames if f in ALGO_BYTES_PER_FRAME else d["launcher_bytes"]
        gbs = algo / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
        tfs = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
        out[f] = {"ms": round(d["ms"], 4), "launches": d["launches"], "kernels": d["kernels"], "algorithmic_bytes": algo,
                  "launcher_algorithmic_bytes": d["launcher_bytes"], "moved_bytes": d["moved_bytes"],
                  "traffic_ratio": d["moved_bytes"] / algo if algo else None,
                  "hbm_gbs": gbs, "hbm_frac": gbs / hbm, "algorithmic_tflops": tfs, "tensor_frac": tfs / tens}
    top = max(out, key=lambda f: out[f]["ms"])
    t = out[top]
    total_algo = sum(ALGO_BYTES_PER_FRAME.values()) * frames
    total_flops = sum(d["flops"] for d in fam.values())
    whole_gbs = total_algo / (ms_step * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": t["hbm_gbs"], "peak": hbm, "unit": "GB/s", "frac": t["hbm_frac"], "traffic": None,
            "kernel": max(t["kernels"], key=t["kernels"].get), "family": top, "family_ms_per_step": t["ms"],
            "algorithmic_bytes": t["algorithmic_bytes"], "moved_bytes": t["moved_bytes"], "traffic_ratio": t["traffic_ratio"],
            "tensor": {"achieved": t["algorithmic_tflops"], "peak": tens, "unit": "TFLOP/s", "frac": t["tensor_frac"],
                       "note": "algorithmic FLOPs against the measured sustained bf16 rate (fp16 operands run at the bf16 rate); "
                               "the 3-pass layers execute 3x their algorithmic FLOPs"},
            "whole_step": {"algorithmic_bytes": total_algo, "ms": ms_step, "achieved": whole_gbs, "frac": whole_gbs / hbm,
                           "algorithmic_tflops": total_flops / (ms_step * 1e-3) / 1e12,
                           "tensor_frac": total_flops / (ms_step * 1e-3) / 1e12 / tens},
            "families": out, "peak_source": peaks["source"] + " (MEASURED_PEAKS.json: HBM copy GB/s, sustained bf16 TFLOP/s)",
            "definition": "achieved = SURVEY 8(d) algorithmic bytes of the family (inputs once + output once per fused layer, fp32) / "
                          "its CUDA-event time in one profiled step; moved_bytes = what the launchers actually request"}
    # measured DRAM bytes of the family's launches in one step (ncu --set full of this build, tools/summarize_ncu.py traffic)
    try:
        import glob
        tr = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))[-1]))
        ent = tr.get(top)
        if ent and "dram_bytes_per_step" in ent:
            roof["traffic"] = ent["dram_bytes_per_step"]
            roof["traffic_note"] = "%s; %s" % (tr.get("_source", ""), ent.get("note", ""))
    except Exception:
        pass
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=6.0)
    ap.add_argument("--ref-batch", type=int, default=8, help="bounded CPU sample: utterances per CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=256, help="concurrent causal streams of the latency measurement")
    ap.add_argument("--stream-steps", type=int, default=1000, help="timed 10 ms hops of the latency measurement (0 = skip)")
    ap.add_argument("--single-stream", action="store_true", help="replay the step graph on one stream instead of alternating two")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every step's launches from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-postnet", action="store_true", help="skip the EaBNet + GaGNet post-filter measurement")
    ap.add_argument("--no-config4", action="store_true", help="skip the 2048-utterance dataset-scale measurement (BASELINE configs[3])")
    ap.add_argument("--no-single", action="store_true", help="skip the one-utterance latency measurement (BASELINE configs[0] shape)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the torch-eager-on-GPU baseline (oracle port on cuda)")
    ap.add_argument("--no-config5", action="store_true", help="skip the training-step object (BASELINE configs[4])")
    ap.add_argument("--train-batch", type=int, default=16, help="utterances (4 s) per GPU of the training-step object")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from eabnet_b200 import EaBNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)

    B, L = args.batch, int(args.seconds * SR)
    M = 9
    # deterministic synthetic weights / waveforms without touching oracle/: uniform init + band-limited noise
    torch.manual_seed(1234 + rank)
    net = EaBNet().eval().to(dev)
    g = torch.Generator().manual_seed(99 + rank)
    wave_host = (0.1 * torch.randn(B, M, L, generator=g)).pin_memory()
    wave = wave_host.to(dev)
    out_host = torch.empty(B, 160 * (L // 160), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    sampler.start()                      # nvidia-smi needs ~1 s to produce its first sample: start before the warm-up
    with torch.no_grad():
        # the device-resident step is replayed from a CUDA graph (EaBNet.graphed_enhance): the ~260 launches of a step are
        # enqueued by the driver, not by this Python thread, so a busy host cannot turn the measurement launch-bound
        # Two graphs (a workspace each) alternate on two streams, like the two compute streams of the host front door: while one
        # step's LSTM holds 108 of the 148 SMs, the other step's kernels use the rest.  Every step is a full pass over the batch.
        if args.no_graph:
            graphs, streams = [], [torch.cuda.current_stream(dev)]
            run_step = lambda i: net.enhance(wave)            # noqa: E731
        else:
            ng = 1 if args.single_stream else int(os.environ.get("EAB_BENCH_STREAMS", "2"))
            graphs = [net.graphed_enhance(wave, private_workspace=True) for _ in range(ng)]
            streams = [torch.cuda.Stream(dev) for _ in range(ng)]

            def run_step(i):
                with torch.cuda.stream(streams[i % ng]):
                    return graphs[i % ng].step()
        cur = torch.cuda.current_stream(dev)

        def fork():                                           # the side streams start after everything queued so far
            if graphs:
                for s_ in streams:
                    s_.wait_stream(cur)

        def join():                                           # ... and the current stream continues after them
            if graphs:
                for s_ in streams:
                    cur.wait_stream(s_)
        fork()
        for i in range(args.warmup):
            run_step(i)
        join()
        barrier()
        sampler.rows.clear()             # keep only samples taken from here on (timed regions)
        # ---- device-resident throughput
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        e0.record()
        fork()
        for i in range(args.steps):
            y = run_step(i)
            marks[i].record(streams[i % len(streams)])
        join()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / args.steps
        done = sorted(e0.elapsed_time(mk) for mk in marks)    # completion times; consecutive differences = per-step pace
        step_ms = [round(b - a, 3) for a, b in zip([0.0] + done[:-1], done)]
        launches = graphs[0].launches if graphs else net.last_launch_count()
        # ---- end to end with host buffers: the dataset-scale public call (eab_enhance_host_batches), `steps` batches,
        # every batch uploaded from pinned host memory and its enhanced audio downloaded inside the timed region
        # (uploads / downloads of neighbouring batches overlap compute on the library's copy streams)
        wave_host2 = wave_host.roll(1, 0).pin_memory()
        out_host2 = torch.empty_like(out_host).pin_memory()
        ins = [wave_host if i % 2 == 0 else wave_host2 for i in range(args.steps)]
        outs = [out_host if i % 2 == 0 else out_host2 for i in range(args.steps)]
        net.enhance_host_batches(ins[:2], outs[:2], dev)
        barrier()
        t0 = time.perf_counter()
        net.enhance_host_batches(ins, outs, dev)
        torch.cuda.synchronize(dev)
        ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
        barrier()
        # ---- the same call on the 16-bit PCM wire format (what the wav files hold): half the H2D / D2H bytes
        pcm_host = [(w * 32768.0).round().clamp(-32768, 32767).to(torch.int16).pin_memory() for w in (wave_host, wave_host2)]
        pcm_out = [torch.empty(out_host.shape, dtype=torch.int16).pin_memory() for _ in range(2)]
        ins16 = [pcm_host[i % 2] for i in range(args.steps)]
        outs16 = [pcm_out[i % 2] for i in range(args.steps)]
        net.enhance_host_batches(ins16[:2], outs16[:2], dev)
        barrier()
        t0 = time.perf_counter()
        net.enhance_host_batches(ins16, outs16, dev)
        torch.cuda.synchronize(dev)
        ms_e2e16 = (time.perf_counter() - t0) * 1e3 / args.steps
        barrier()
        # single-call latency form (one batch, nothing to overlap with)
        t0 = time.perf_counter()
        net.enhance_host(wave_host, out_host, dev)
        torch.cuda.synchronize(dev)
        ms_e2e_single = (time.perf_counter() - t0) * 1e3
        assert torch.isfinite(out_host2).all()
        # ---- per-kernel-family timing of one more step (CUDA events around every launch)
        net.profile(True)
        net.enhance(wave)
        prof = net.profile_summary()
        net.profile(False)
        clocks = sampler.stop()
        assert torch.isfinite(y).all()

    def optional(fn, *a):
        """the secondary objects must never cost the headline line: a failure is reported inside the object"""
        try:
            return fn(*a)
        except Exception as e:                       # noqa: BLE001
            return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    c4 = optional(config4, args, dev, net, world, rank, [wave_host, wave_host2], [out_host, out_host2]) if not args.no_config4 else None
    single = optional(single_utterance, dev, net) if (rank == 0 and not args.no_single) else None
    eager = optional(gpu_eager_baseline, args, dev) if (rank == 0 and world == 1 and not args.no_gpu_eager) else None
    latency = optional(stream_latency, args, dev) if args.stream_steps > 0 else None
    postnet = optional(postnet_throughput, args, dev, wave) if (rank == 0 and not args.no_postnet) else None
    del graphs
    torch.cuda.empty_cache()
    c5 = optional(config5_training, args, dev, world, rank, local) if not args.no_config5 else None

    from eabnet_b200.shard import max_over_ranks
    ms, ms_e2e, ms_e2e16 = max_over_ranks([ms, ms_e2e, ms_e2e16], dev)
    audio_s = world * B * args.seconds
    value = audio_s / (ms * 1e-3)
    e2e = audio_s / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = load_peaks()
        roof = build_roofline(prof, ms, B * (1 + L // 160), peaks)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "batch_per_gpu": B, "seconds": args.seconds, "frames": 1 + L // 160,
                           "parallelism": "utterance shards, %d rank(s), no collective" % world,
                           "l2": "inputs (%.0f MB/step) and activations exceed the 126 MB L2" % (B * M * L * 4 / 1e6),
                           "launch": "python enqueue" if args.no_graph else ("CUDA graph replay of eab_enhance" + ("" if args.single_stream else
                                     ", steps alternating on two streams (two graphs, a workspace each)"))},
                "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": B * M * L * 4,
                        "d2h_bytes_per_step": B * 160 * (L // 160) * 4,
                        "api": "EaBNet.enhance_host_batches (eab_enhance_host_batches): %d host batches per call, wall clock "
                               "around the call, copies overlapped with compute" % args.steps,
                        "single_batch_call_ms": ms_e2e_single},
                "e2e_pcm16": {"value": audio_s / (ms_e2e16 * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e16, "h2d_bytes_per_step": B * M * L * 2,
                              "d2h_bytes_per_step": B * 160 * (L // 160) * 2,
                              "api": "EaBNet.enhance_host_batches on int16 PCM batches (eab_enhance_host_batches_pcm16): conversions fused into "
                                     "the STFT staging and the iSTFT store"},
                "gpu_launches": launches * args.steps, "step_ms": step_ms,
                "roofline": roof, "clocks": clocks, "kernels": prof}
        if c4 is not None:
            line["config4"] = c4
        if single is not None:
            line["single_utterance"] = single
        if eager is not None:
            line["gpu_eager_baseline"] = eager
        if latency is not None:
            line["latency"] = latency
        if postnet is not None:
            line["postnet"] = postnet
        if c5 is not None:
            line["config5"] = c5
        if world == 1 and not args.no_cpu_baseline:
            val, med, cores, sample = cpu_reference_throughput(args.ref_batch, args.seconds, 3, 1)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
