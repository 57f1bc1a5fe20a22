"""Headline benchmark: clips/sec of the 16-frame TimeSformer-B + DistilBERT + EgoNCE training step.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (the reference algorithm on the host CPU cores)

One "step" = zero_grad -> FrozenInTime forward (text + video towers) -> ONE packed embedding/tag all-gather ->
fused similarity + EgoNCE -> backward (every dgrad/wgrad) -> DDP gradient all-reduce -> AdamW step, on a
synthetic batch of BASELINE.json's shape: per-GPU batch 64 clips of 16 x 3 x 224 x 224 fp32 + 16-token text,
G = 64*N (weak scaling).  Prints ONE JSON line (contract in the task statement / DESIGN.md section 7).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

# algorithmic FLOPs (multiply-add = 2)
def flops_per_clip(T, L, N=196, D=768, H=12, HID=3072, depth=12, text_layers=6):
    S = 1 + T * N
    blk = S * (2 * 2 * D * 3 * D + 2 * 2 * D * D + 2 * 2 * D * HID) + H * N * 4 * 64 * T * (T + 1) + \
        H * T * 4 * 64 * N * (N + 1) + 2 * H * 4 * 64 * S
    patch = 2 * T * N * D * D
    video = depth * blk + patch + 2 * D * 256
    text = text_layers * (L * (8 * D * D + 4 * D * HID) + H * 4 * 64 * L * L)
    return 3 * (video + text) - patch, video, text


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def host_cores():
    """CPU threads this process may really use: affinity mask capped by the cgroup CPU quota (os.cpu_count() reports
    the host's cores even inside a CPU-limited container, and oversubscribing torch threads is catastrophic)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0:
                n = min(n, max(1, quota // period))
        except (OSError, ValueError):
            pass
    return max(1, min(n, 32))   # beyond ~32 threads the fp32 eager path at this size scales negatively


class ClockSampler:
    """nvidia-smi SM clock / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step_rate(T, L, steps, warmup, batch=2, seed=0):
    """The reference algorithm (oracle port, torch fp32, all host threads) on a bounded sample of the workload:
    `batch` clips of T frames, fwd + bwd + AdamW.  Returns clips/s and a description."""
    from oracle import reference_port as rp
    from egovlp_b200 import synthetic as syn
    cores = host_cores()
    torch.set_num_threads(cores)
    dims = syn.model_dims(num_frames=max(T, 4))
    params = {k: v.requires_grad_(True) for k, v in syn.seeded_state_dict(dims, seed=seed).items()}
    opt = torch.optim.AdamW(list(params.values()), lr=3e-5, eps=1e-6, weight_decay=0.0)
    data = {"video": syn.synthetic_video(batch, T, seed=seed), "text": syn.synthetic_text(batch, L, seed=seed)}
    verb, noun = syn.synthetic_tags(batch, seed=seed)

    def step():
        opt.zero_grad(set_to_none=True)
        t, v = rp.frozen_in_time_forward(data, params)
        loss = rp.egonce_loss(rp.sim_matrix(t, v), rp.sim_matrix(verb, verb), rp.sim_matrix(noun, noun))
        loss.backward()
        opt.step()
        return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    while done < steps and (done == 0 or time.perf_counter() - t0 < 45.0):     # bounded: ~10-45 s of CPU work
        step()
        done += 1
    dt = time.perf_counter() - t0
    return (batch * done / dt, dt / done, cores,
            f"oracle port, fp32, {batch} clips x {T}f x 224^2 + {L} tokens, fwd+bwd+AdamW, {done} timed steps on {cores} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 4))
    warm = max(1, min(args.warmup, 1))
    value, s_per_step, cores, sample = cpu_reference_step_rate(args.frames, args.text_len, steps, warm)
    line = {"impl": "reference", "metric": "clips/sec, 16-frame TimeSformer-B + DistilBERT + EgoNCE training step",
            "value": value, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (seeded video/text/tags, seeded random-init weights)",
            "config": workload_config(args, 1, cpu_sample=True),
            "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world, cpu_sample=False):
    return {"workload": f"EgoClip pretraining step: TimeSformer-B {args.frames}f x 224^2 p16 (divided space-time attention) "
                        f"+ DistilBERT L={args.text_len} + EgoNCE, per-GPU batch {args.batch}"
                        + (" [CPU arm: bounded sample of 2 clips per step]" if cpu_sample else ""),
            "global_batch": args.batch * world, "frames": args.frames, "text_len": args.text_len,
            "parallelism": f"dp{world}", "optimizer": "AdamW (HF semantics) lr 3e-5",
            "l2_policy": "per-step working set (~90 GB of activations) >> 126 MB L2; no explicit flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--text-len", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from egovlp_b200 import _lib, ops, synthetic as syn
    from egovlp_b200.distributed import egoclip_step_loss
    from egovlp_b200.model.loss import EgoNCE
    from egovlp_b200.model.model import FrozenInTime
    from egovlp_b200.optim import AdamW

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert args.warmup >= 3 or args.steps <= 2, "use at least 3 warm-up steps for a reportable number"

    B, T, L = args.batch, args.frames, args.text_len
    net = FrozenInTime({"model": "SpaceTimeTransformer", "arch_config": "base_patch16_224", "num_frames": max(T, 4),
                        "pretrained": True, "time_init": "zeros"},
                       {"model": "distilbert-base-uncased", "pretrained": True, "input": "text"})
    net.load_state_dict(syn.seeded_state_dict(syn.model_dims(num_frames=max(T, 4)), seed=0), strict=True)
    net.to(dev)
    model = net
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], gradient_as_bucket_view=True)
    loss_fn = EgoNCE()
    opt = AdamW(net.parameters(), lr=3e-5)

    seed = 1000 * 0 + rank                                   # identical weights, per-rank data
    host = {"video": syn.synthetic_video(B, T, seed=seed).pin_memory(),
            "ids": syn.synthetic_text(B, L, seed=seed)["input_ids"].pin_memory(),
            "mask": syn.synthetic_text(B, L, seed=seed)["attention_mask"].pin_memory()}
    verb_h, noun_h = syn.synthetic_tags(B, seed=seed)
    host["verb"], host["noun"] = verb_h.pin_memory(), noun_h.pin_memory()

    def to_device():
        return {"video": host["video"].to(dev, non_blocking=True),
                "text": {"input_ids": host["ids"].to(dev, non_blocking=True),
                         "attention_mask": host["mask"].to(dev, non_blocking=True)},
                "verb_vec": host["verb"].to(dev, non_blocking=True), "noun_vec": host["noun"].to(dev, non_blocking=True)}

    resident = to_device()
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())

    def step(data):
        opt.zero_grad(set_to_none=True)
        loss = egoclip_step_loss(model, loss_fn, data)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(n):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), out

    for _ in range(args.warmup):
        loss = step(resident)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    ops.profile_gemm(True)
    ms_total, loss = timed(args.steps, lambda: step(resident))
    gemm_flops, gemm_ms, gemm_calls = ops.profile_gemm(False)
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(loss.item())
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total / 1e3)

    e2e = None
    if not args.no_e2e:
        from egovlp_b200.data import DevicePrefetcher

        def host_batches(n, video):
            for _ in range(n):
                yield {"video": video, "text": {"input_ids": host["ids"], "attention_mask": host["mask"]},
                       "verb_vec": host["verb"], "noun_vec": host["noun"]}

        def run_e2e(video):
            n = args.steps + 1
            it = iter(DevicePrefetcher(host_batches(n, video), dev))     # every batch: pinned host -> device copy
            step(next(it)).item()                                        # untimed first step
            def one():
                return step(next(it)).item()                             # D2H read of the loss every step
            ms, _ = timed(args.steps, one)
            return ms

        ms_e2e = run_e2e(host["video"])
        e2e = {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": "clips/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
               "note": "per step: pinned-host fp32 video/ids/mask/tags -> device (one batch in flight on a copy "
                       "stream, egovlp_b200.data.DevicePrefetcher), model(data) public API, loss.item()"}
        # same step fed with uint8 frames (normalisation fused into the patch-embedding kernel): 4x fewer H2D bytes
        mean = torch.tensor(syn.IMAGENET_MEAN).view(1, 1, 3, 1, 1)
        std = torch.tensor(syn.IMAGENET_STD).view(1, 1, 3, 1, 1)
        video_u8 = ((host["video"] * std + mean).clamp(0, 1) * 255).round().to(torch.uint8).pin_memory()
        ms_u8 = run_e2e(video_u8)
        e2e["uint8_frames"] = {"value": B * world * args.steps / (ms_u8 / 1e3), "ms_per_step": ms_u8 / args.steps,
                               "h2d_bytes_per_step": h2d_bytes - host["video"].numel() * 3}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        f_step, f_video, f_text = flops_per_clip(T, L)
        peak_tf = peaks["bf16_tflops_sustained"]
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        # DRAM bytes of the same launches from the committed ncu capture (valid for the default workload only)
        traffic, traffic_src = None, None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_gemm_dram_traffic.json")
        if os.path.exists(tpath) and (args.batch, T, L) == (64, 16, 16):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_step"] / tj["launches_per_step"]
            traffic_src = "profiles/r1_gemm_dram_traffic.json (ncu dram__bytes_read+write.sum, mean per GEMM launch of one step)"
        line = {"metric": "clips/sec, 16-frame TimeSformer-B + DistilBERT + EgoNCE training step",
                "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic (seeded video/text/tags, seeded random-init weights; text-tower dropout 0.1 active as in the reference's train mode)",
                "config": workload_config(args, world), "loss": loss_val,
                "step_flop_fraction_of_peak": value / world * f_step / (peak_tf * 1e12),
                "gflop_per_clip_step": f_step / 1e9,
                "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (all fwd/dgrad/wgrad launches)",
                             "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": achieved / peak_tf if achieved else None, "traffic": traffic,
                             "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                             "algorithmic_flop_per_launch": gemm_flops / gemm_calls if gemm_calls else None,
                             "launches_per_step": gemm_calls / args.steps, "share_of_step": gemm_ms / ms_total,
                             "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches}
        if not args.no_cpu_baseline and world == 1:
            v, s, cores, sample = cpu_reference_step_rate(T, L, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
